// Minimal stand-in for AviSynth+'s public header `avisynth.h`.
//
// PURPOSE: test infrastructure. Neither the reference (/root/reference) nor this repo ships
// AviSynth+'s header and it is not installed in the build image, so this file provides the
// small subset of the AviSynth+ plugin API (names and call shapes as documented by AviSynth+)
// that (a) the unmodified reference sources need in order to compile for the oracle build
// (oracle/Makefile -> oracle/_ref/) and (b) our own plugin shim (host/sangnom2_plugin.cpp)
// uses. The shim only calls methods that exist with the same meaning in the real header, so
// it also builds against a real AviSynth+ SDK (CMake: -DAVS_INCLUDE_DIR=<dir>; make: AVS_INC=<dir>).
//
// This is NOT a reimplementation of AviSynth: there is no script parser, no cache, no audio.
// The matching fake host (../fake_host.cpp) implements IScriptEnvironment just far enough
// to load a plugin through AvisynthPluginInit3, call a registered factory with an AVSValue
// array and pull frames.
//
// Use sites this subset was derived from: /root/reference/src/SangNom2.h:40-67 and
// /root/reference/src/SangNom2.cpp:275-484 (ctor, GetFrame, factories, plugin init).
#pragma once

#include <algorithm>
#include <cstdarg>
#include <cstddef>
#include <cstdint>
#include <mutex>
#include <vector>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#ifndef _MSC_VER
#ifndef __stdcall
#define __stdcall
#endif
#ifndef __cdecl
#define __cdecl
#endif
#ifndef __declspec
#define __declspec(x)
#endif
#endif

#ifndef AVS_FORCEINLINE
#if defined(_MSC_VER)
#define AVS_FORCEINLINE __forceinline
#else
#define AVS_FORCEINLINE inline __attribute__((always_inline))
#endif
#endif

#define AVS_STUB_HEADER 1

typedef unsigned char BYTE;

enum { AVISYNTH_INTERFACE_VERSION = 8 };
enum { FRAME_ALIGN = 64 };

// Plane selectors (same symbolic names as AviSynth+).
enum AvsPlane {
    PLANAR_Y = 1 << 0,
    PLANAR_U = 1 << 1,
    PLANAR_V = 1 << 2,
    PLANAR_ALIGNED = 1 << 3,
    PLANAR_A = 1 << 4,
    PLANAR_R = 1 << 5,
    PLANAR_G = 1 << 6,
    PLANAR_B = 1 << 7,
};

// CPU flags (only the one the reference tests).
enum { CPUF_SSE2 = 0x20 };

// Cache hints / MT modes (only what filters answer).
enum CachePolicyHint {
    CACHE_GET_MTMODE = 509,
};
enum MtMode {
    MT_INVALID = 0,
    MT_NICE_FILTER = 1,
    MT_MULTI_INSTANCE = 2,
    MT_SERIALIZED = 3,
};

struct AVS_Linkage { int Size; };
extern const AVS_Linkage* AVS_linkage;

// ---------------------------------------------------------------------------------------------
// VideoInfo. The real header encodes the format in `pixel_type` bit fields; the stub keeps the
// decoded description in plain members and answers the same query methods.
struct VideoInfo {
    int width = 0, height = 0;
    unsigned fps_numerator = 25, fps_denominator = 1;
    int num_frames = 0;

    // stub-only description of the colour format
    int stub_components = 1;      // 1 = Y, 3 = YUV, 4 = YUVA
    int stub_sub_w = 0;           // log2 horizontal chroma subsampling
    int stub_sub_h = 0;           // log2 vertical chroma subsampling
    int stub_bits = 8;            // 8,10,12,14,16 or 32 (float)
    bool stub_rgb = false;
    bool stub_planar = true;

    bool HasVideo() const { return width != 0; }
    bool IsRGB() const { return stub_rgb; }
    bool IsPlanar() const { return stub_planar; }
    bool IsYUV() const { return !stub_rgb && stub_components == 3; }
    bool IsYUVA() const { return !stub_rgb && stub_components == 4; }
    bool IsY() const { return !stub_rgb && stub_components == 1; }
    bool Is420() const { return !stub_rgb && stub_components >= 3 && stub_sub_w == 1 && stub_sub_h == 1; }
    bool Is422() const { return !stub_rgb && stub_components >= 3 && stub_sub_w == 1 && stub_sub_h == 0; }
    bool Is444() const { return !stub_rgb && stub_components >= 3 && stub_sub_w == 0 && stub_sub_h == 0; }
    bool IsYV411() const { return !stub_rgb && stub_components >= 3 && stub_sub_w == 2 && stub_sub_h == 0; }
    int NumComponents() const { return stub_components; }
    int ComponentSize() const { return stub_bits <= 8 ? 1 : (stub_bits <= 16 ? 2 : 4); }
    int BitsPerComponent() const { return stub_bits; }
    int GetPlaneWidthSubsampling(int plane) const { return (plane == PLANAR_U || plane == PLANAR_V) ? stub_sub_w : 0; }
    int GetPlaneHeightSubsampling(int plane) const { return (plane == PLANAR_U || plane == PLANAR_V) ? stub_sub_h : 0; }
    int RowSize(int plane = 0) const { return (width >> GetPlaneWidthSubsampling(plane)) * ComponentSize(); }
};

// ---------------------------------------------------------------------------------------------
// Frames. One heap block (VideoFrameBuffer) per frame holding all planes, like AviSynth+; pitch rounded up to the
// requested alignment. Frame buffers are recycled through a small free list, like AviSynth+'s frame registry: a frame
// server does not go to the system allocator (and fault in fresh pages) for every output frame.
class StubFramePool {
    struct Entry { void* p; size_t bytes, align; };
    static std::mutex& mu() { static std::mutex m; return m; }
    static std::vector<Entry>& list() { static std::vector<Entry> v; return v; }
public:
    static void* take(size_t bytes, size_t align)
    {
        {
            std::lock_guard<std::mutex> lk(mu());
            auto& v = list();
            for (size_t i = 0; i < v.size(); ++i)
                if (v[i].bytes == bytes && v[i].align == align) { void* p = v[i].p; v[i] = v.back(); v.pop_back(); return p; }
        }
        void* mem = nullptr;
        if (::posix_memalign(&mem, align, bytes ? bytes : 1) != 0) throw std::bad_alloc();
        return mem;
    }
    static void give(void* p, size_t bytes, size_t align)
    {
        std::lock_guard<std::mutex> lk(mu());
        auto& v = list();
        if (v.size() < 4096) v.push_back(Entry{ p, bytes, align }); else std::free(p);
    }
};

// The block of memory behind a frame (AviSynth+: VideoFrame::GetFrameBuffer()).
class VideoFrameBuffer {
    friend class VideoFrame;
    BYTE* data = nullptr;
    int data_size = 0;
public:
    const BYTE* GetReadPtr() const { return data; }
    BYTE* GetWritePtr() { return data; }
    int GetDataSize() const { return data_size; }
};

class VideoFrame {
    friend class PVideoFrame;
    friend class FakeHostAccess;
    int refcount = 0;
    struct PlaneBuf { BYTE* data = nullptr; int pitch = 0, row_size = 0, height = 0; size_t bytes = 0; };
    VideoFrameBuffer vfb_;
    size_t vfb_bytes_ = 0, vfb_align_ = 0;
    // buffers go back to the pool of the module that created the frame (this header is compiled into several shared
    // objects, each with its own StubFramePool statics; the destructor may run in any of them)
    void (*release_)(void*, size_t, size_t) = nullptr;
    PlaneBuf planes_[4];
    std::vector<std::pair<std::string, int64_t>> props_;

    static int index_of(int plane) {
        switch (plane & ~PLANAR_ALIGNED) {
            case 0: case PLANAR_Y: case PLANAR_R: return 0;
            case PLANAR_U: case PLANAR_G: return 1;
            case PLANAR_V: case PLANAR_B: return 2;
            case PLANAR_A: return 3;
            default: return 0;
        }
    }

public:
    VideoFrame(const VideoInfo& vi, int align, bool poison) {
        const int n = vi.NumComponents();
        static const int ids[4] = { PLANAR_Y, PLANAR_U, PLANAR_V, PLANAR_A };
        if (align < 16) align = 16;
        release_ = &StubFramePool::give;
        size_t total = 0;
        for (int i = 0; i < n; ++i) {
            PlaneBuf& p = planes_[i];
            const int sw = (i == 1 || i == 2) ? vi.GetPlaneWidthSubsampling(ids[i]) : 0;
            const int sh = (i == 1 || i == 2) ? vi.GetPlaneHeightSubsampling(ids[i]) : 0;
            p.row_size = (vi.width >> sw) * vi.ComponentSize();
            p.height = vi.height >> sh;
            p.pitch = (p.row_size + align - 1) / align * align;
            p.bytes = ((size_t)p.pitch * (size_t)std::max(p.height, 1) + (size_t)align - 1) / (size_t)align * (size_t)align;
            total += p.bytes;
        }
        vfb_bytes_ = total ? total : 1;
        vfb_align_ = (size_t)align;
        vfb_.data = static_cast<BYTE*>(StubFramePool::take(vfb_bytes_, vfb_align_));
        vfb_.data_size = (int)vfb_bytes_;
        size_t off = 0;
        for (int i = 0; i < n; ++i) { planes_[i].data = vfb_.data + off; off += planes_[i].bytes; }
        // New frames hold garbage in a real host; poison them so that reads of never-written
        // output (e.g. the alpha plane in the reference) are visible in tests. Without poisoning the
        // memory is left as it is, like a real host's recycled frame buffers.
        if (poison) std::memset(vfb_.data, 0xCD, vfb_bytes_);
    }
    ~VideoFrame() { if (vfb_.data) release_(vfb_.data, vfb_bytes_, vfb_align_); }
    VideoFrameBuffer* GetFrameBuffer() const { return const_cast<VideoFrameBuffer*>(&vfb_); }
    VideoFrame(const VideoFrame&) = delete;
    VideoFrame& operator=(const VideoFrame&) = delete;

    int GetPitch(int plane = 0) const { return planes_[index_of(plane)].pitch; }
    int GetRowSize(int plane = 0) const { return planes_[index_of(plane)].row_size; }
    int GetHeight(int plane = 0) const { return planes_[index_of(plane)].height; }
    const BYTE* GetReadPtr(int plane = 0) const { return planes_[index_of(plane)].data; }
    BYTE* GetWritePtr(int plane = 0) { return planes_[index_of(plane)].data; }
    bool IsWritable() const { return refcount == 1; }

    // frame properties: just enough to observe that NewVideoFrameP copies them
    void stub_set_prop(const char* key, int64_t v) { props_.emplace_back(key, v); }
    bool stub_get_prop(const char* key, int64_t* v) const {
        for (auto& kv : props_) if (kv.first == key) { *v = kv.second; return true; }
        return false;
    }
    void stub_copy_props_from(const VideoFrame& o) { props_ = o.props_; }
};

class PVideoFrame {
    VideoFrame* p = nullptr;
    void set(VideoFrame* x) {
        if (x) ++x->refcount;
        if (p && --p->refcount == 0) delete p;
        p = x;
    }
public:
    PVideoFrame() {}
    PVideoFrame(VideoFrame* x) { set(x); }
    PVideoFrame(const PVideoFrame& o) { set(o.p); }
    PVideoFrame& operator=(const PVideoFrame& o) { set(o.p); return *this; }
    ~PVideoFrame() { set(nullptr); }
    VideoFrame* operator->() const { return p; }
    operator void*() const { return p; }
    bool operator!() const { return !p; }
};

// ---------------------------------------------------------------------------------------------
class IScriptEnvironment;

class IClip {
    friend class PClip;
    friend class AVSValue;
    int refcnt = 0;
public:
    IClip() {}
    virtual ~IClip() {}
    virtual int __stdcall GetVersion() { return AVISYNTH_INTERFACE_VERSION; }
    virtual PVideoFrame __stdcall GetFrame(int n, IScriptEnvironment* env) = 0;
    virtual bool __stdcall GetParity(int n) = 0;
    virtual int __stdcall SetCacheHints(int cachehints, int frame_range) = 0;
    virtual const VideoInfo& __stdcall GetVideoInfo() = 0;
};

class PClip {
    IClip* p = nullptr;
    void set(IClip* x) {
        if (x) ++x->refcnt;
        if (p && --p->refcnt == 0) delete p;
        p = x;
    }
public:
    PClip() {}
    PClip(IClip* x) { set(x); }
    PClip(const PClip& o) { set(o.p); }
    PClip& operator=(const PClip& o) { set(o.p); return *this; }
    ~PClip() { set(nullptr); }
    IClip* operator->() const { return p; }
    IClip* get() const { return p; }
    operator void*() const { return p; }
    bool operator!() const { return !p; }
};

class GenericVideoFilter : public IClip {
protected:
    PClip child;
    VideoInfo vi;
public:
    GenericVideoFilter(PClip _child) : child(_child) { vi = child->GetVideoInfo(); }
    PVideoFrame __stdcall GetFrame(int n, IScriptEnvironment* env) override { return child->GetFrame(n, env); }
    bool __stdcall GetParity(int n) override { return child->GetParity(n); }
    int __stdcall SetCacheHints(int, int) override { return 0; }
    const VideoInfo& __stdcall GetVideoInfo() override { return vi; }
};

// ---------------------------------------------------------------------------------------------
// AVSValue: undefined / clip / bool / int / float / string / array.
class AVSValue {
    char type = 'v';
    PClip clip_;
    bool b_ = false;
    int i_ = 0;
    float f_ = 0.f;
    const char* s_ = nullptr;
    const AVSValue* arr_ = nullptr;
    int arr_n_ = 0;
public:
    AVSValue() {}
    AVSValue(IClip* c) : type('c'), clip_(c) {}
    AVSValue(const PClip& c) : type('c'), clip_(c) {}
    AVSValue(bool b) : type('b'), b_(b) {}
    AVSValue(int i) : type('i'), i_(i) {}
    AVSValue(float f) : type('f'), f_(f) {}
    AVSValue(double f) : type('f'), f_((float)f) {}
    AVSValue(const char* s) : type('s'), s_(s) {}
    AVSValue(const AVSValue* a, int n) : type('a'), arr_(a), arr_n_(n) {}

    bool Defined() const { return type != 'v'; }
    bool IsClip() const { return type == 'c'; }
    bool IsBool() const { return type == 'b'; }
    bool IsInt() const { return type == 'i'; }
    bool IsFloat() const { return type == 'f' || type == 'i'; }
    bool IsString() const { return type == 's'; }
    bool IsArray() const { return type == 'a'; }

    PClip AsClip() const { return clip_; }
    bool AsBool() const { return b_; }
    int AsInt() const { return i_; }
    const char* AsString() const { return s_; }
    double AsFloat() const { return type == 'i' ? (double)i_ : (double)f_; }
    bool AsBool(bool def) const { return Defined() ? b_ : def; }
    int AsInt(int def) const { return Defined() ? i_ : def; }
    double AsFloat(float def) const { return Defined() ? AsFloat() : (double)def; }
    const char* AsString(const char* def) const { return Defined() ? s_ : def; }
    int ArraySize() const { return type == 'a' ? arr_n_ : 1; }
    // Out-of-range subscripts on a real AVSValue array are a host-side assertion; AviSynth+'s
    // release build returns the value itself. Return an undefined value instead so that
    // `AsInt(def)` / `AsBool(def)` yield `def`.
    const AVSValue& operator[](int index) const {
        static const AVSValue undefined;
        if (type == 'a' && index >= 0 && index < arr_n_) return arr_[index];
        return undefined;
    }
};

// ---------------------------------------------------------------------------------------------
class AvisynthError {
public:
    const char* const msg;
    AvisynthError(const char* m) : msg(m) {}
};

class IScriptEnvironment {
public:
    typedef AVSValue(__cdecl* ApplyFunc)(AVSValue args, void* user_data, IScriptEnvironment* env);

    virtual ~IScriptEnvironment() {}
    virtual int __stdcall GetCPUFlags() = 0;
    virtual char* __stdcall SaveString(const char* s, int length = -1) = 0;
    virtual void __stdcall ThrowError(const char* fmt, ...) = 0;   // never returns (throws AvisynthError)
    virtual void __stdcall AddFunction(const char* name, const char* params, ApplyFunc apply, void* user_data) = 0;
    virtual bool __stdcall FunctionExists(const char* name) = 0;
    virtual PVideoFrame __stdcall NewVideoFrame(const VideoInfo& vi, int align = FRAME_ALIGN) = 0;
    virtual PVideoFrame __stdcall NewVideoFrameP(const VideoInfo& vi, PVideoFrame* prop_src, int align = FRAME_ALIGN) = 0;
    virtual void __stdcall BitBlt(BYTE* dstp, int dst_pitch, const BYTE* srcp, int src_pitch, int row_size, int height) = 0;
};
