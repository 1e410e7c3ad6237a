// Fake AviSynth host: test infrastructure (no part of the product path).
//
// Implements just enough of IScriptEnvironment (see avs_stub/avisynth.h) to do what a real
// frameserver does with a plugin: dlopen it, call AvisynthPluginInit3, look up a registered
// function, call its factory with an AVSValue argument array built from named arguments, and
// pull frames from the returned clip. The same host drives BOTH the unmodified reference
// plugin (built by oracle/Makefile into oracle/_ref/) and our plugin, so the plugin-level
// parity tests run the identical call sequence against each (tests/test_plugin_*.py).
//
// Exposed as a plain C API (fh_*) for ctypes.
#include "avisynth.h"

#include <dlfcn.h>

#include <map>
#include <memory>
#include <mutex>

const AVS_Linkage* AVS_linkage = nullptr;   // the host's own copy; plugins carry theirs

namespace {

struct RegisteredFunction {
    std::string name, params;
    IScriptEnvironment::ApplyFunc apply;
    void* user_data;
    // parsed signature
    std::vector<char> types;            // 'c','i','b','f','s'
    std::vector<std::string> names;     // "" for unnamed
};

static void parse_params(RegisteredFunction& f) {
    const std::string& p = f.params;
    size_t i = 0;
    while (i < p.size()) {
        std::string name;
        if (p[i] == '[') {
            size_t j = p.find(']', i);
            name = p.substr(i + 1, j - i - 1);
            i = j + 1;
        }
        if (i >= p.size()) break;
        f.types.push_back(p[i]);
        f.names.push_back(name);
        ++i;
        while (i < p.size() && (p[i] == '+' || p[i] == '*')) ++i;
    }
}

class FakeEnv : public IScriptEnvironment {
public:
    int cpu_flags = CPUF_SSE2;
    bool has_v8 = true;
    bool poison_new_frames = true;
    std::vector<RegisteredFunction> functions;
    std::vector<std::unique_ptr<char[]>> strings;
    std::vector<void*> dl_handles;
    long frames_allocated = 0;
    char errbuf[1024];

    ~FakeEnv() override {}

    int __stdcall GetCPUFlags() override { return cpu_flags; }
    char* __stdcall SaveString(const char* s, int length = -1) override {
        size_t n = length < 0 ? std::strlen(s) : (size_t)length;
        strings.emplace_back(new char[n + 1]);
        std::memcpy(strings.back().get(), s, n);
        strings.back()[n] = 0;
        return strings.back().get();
    }
    void __stdcall ThrowError(const char* fmt, ...) override {
        va_list ap;
        va_start(ap, fmt);
        std::vsnprintf(errbuf, sizeof errbuf, fmt, ap);
        va_end(ap);
        throw AvisynthError(SaveString(errbuf));
    }
    void __stdcall AddFunction(const char* name, const char* params, ApplyFunc apply, void* user_data) override {
        RegisteredFunction f{ name, params, apply, user_data, {}, {} };
        parse_params(f);
        functions.push_back(f);
    }
    bool __stdcall FunctionExists(const char* name) override {
        if (std::strcmp(name, "propShow") == 0) return has_v8;   // the v8 probe plugins use
        for (auto& f : functions) if (f.name == name) return true;
        return false;
    }
    PVideoFrame __stdcall NewVideoFrame(const VideoInfo& vi, int align = FRAME_ALIGN) override {
        ++frames_allocated;
        // like AviSynth+, the requested alignment is a minimum: frames never get less than
        // FRAME_ALIGN, so source and filter-allocated frames of one format share a pitch (the
        // reference's whole-plane memcpy for disabled planes relies on that, SangNom2.cpp:372)
        return PVideoFrame(new VideoFrame(vi, std::max(align, (int)FRAME_ALIGN), poison_new_frames));
    }
    PVideoFrame __stdcall NewVideoFrameP(const VideoInfo& vi, PVideoFrame* prop_src, int align = FRAME_ALIGN) override {
        PVideoFrame f = NewVideoFrame(vi, align);
        if (prop_src && *prop_src) f->stub_copy_props_from(*(*prop_src).operator->());
        return f;
    }
    void __stdcall BitBlt(BYTE* dstp, int dst_pitch, const BYTE* srcp, int src_pitch, int row_size, int height) override {
        for (int y = 0; y < height; ++y)
            std::memcpy(dstp + (int64_t)y * dst_pitch, srcp + (int64_t)y * src_pitch, (size_t)row_size);
    }
};

// A source clip whose frames are filled from the test (numpy) side.
class SourceClip : public IClip {
public:
    VideoInfo vi;
    int parity_mode;   // 0: always false, 1: always true, 2: alternate, true for even n
    std::vector<PVideoFrame> frames;
    std::vector<int> requests;   // log of GetFrame(n) calls, for the batching tests
    std::mutex mu;

    bool log_requests = true;
    int fail_at = -1;            // fault injection: GetFrame(fail_at) raises a script error
    // stored < num_frames: a long clip that repeats its `stored` frames (for throughput runs)
    SourceClip(const VideoInfo& v, int pm, int stored = 0) : vi(v), parity_mode(pm) {
        frames.resize((size_t)(stored > 0 ? stored : v.num_frames));
        for (auto& f : frames) f = PVideoFrame(new VideoFrame(vi, FRAME_ALIGN, false));
        log_requests = stored <= 0;
    }
    PVideoFrame __stdcall GetFrame(int n, IScriptEnvironment* env) override {
        std::lock_guard<std::mutex> lk(mu);
        if (log_requests) requests.push_back(n);
        if (n == fail_at) env->ThrowError("FakeSource: injected failure at frame %d", n);
        n = std::max(0, std::min(n, vi.num_frames - 1));
        return frames[(size_t)n % frames.size()];
    }
    bool __stdcall GetParity(int n) override {
        switch (parity_mode) {
            case 0: return false;
            case 1: return true;
            default: return (n & 1) == 0;
        }
    }
    int __stdcall SetCacheHints(int, int) override { return 0; }
    const VideoInfo& __stdcall GetVideoInfo() override { return vi; }
};

struct ClipHandle { PClip clip; SourceClip* source = nullptr; };
struct FrameHandle { PVideoFrame frame; };

static const int kPlaneIds[4] = { PLANAR_Y, PLANAR_U, PLANAR_V, PLANAR_A };

static void set_err(char* err, int errlen, const char* msg) {
    if (err && errlen > 0) { std::snprintf(err, (size_t)errlen, "%s", msg); }
}

}  // namespace

extern "C" {

void* fh_env_create(int cpu_flags, int has_v8, int poison_new_frames) {
    auto* e = new FakeEnv();
    e->cpu_flags = cpu_flags;
    e->has_v8 = has_v8 != 0;
    e->poison_new_frames = poison_new_frames != 0;
    return e;
}

void fh_env_destroy(void* env) {
    auto* e = static_cast<FakeEnv*>(env);
    // plugin code must outlive every clip created from it; handles are closed last
    std::vector<void*> handles = e->dl_handles;
    delete e;
    for (void* h : handles) dlclose(h);
}

// Returns the plugin's name string (AvisynthPluginInit3's result) or NULL with err filled.
const char* fh_load_plugin(void* env, const char* path, char* err, int errlen) {
    auto* e = static_cast<FakeEnv*>(env);
    void* h = dlopen(path, RTLD_NOW | RTLD_LOCAL);
    if (!h) { set_err(err, errlen, dlerror()); return nullptr; }
    typedef const char* (*InitFn)(IScriptEnvironment*, const AVS_Linkage*);
    auto init = reinterpret_cast<InitFn>(dlsym(h, "AvisynthPluginInit3"));
    if (!init) { set_err(err, errlen, "AvisynthPluginInit3 not exported"); dlclose(h); return nullptr; }
    static const AVS_Linkage linkage{ (int)sizeof(AVS_Linkage) };
    e->dl_handles.push_back(h);
    try {
        return init(e, &linkage);
    } catch (const AvisynthError& x) {
        set_err(err, errlen, x.msg);
        return nullptr;
    }
}

int fh_function_count(void* env) { return (int)static_cast<FakeEnv*>(env)->functions.size(); }
const char* fh_function_name(void* env, int i) { return static_cast<FakeEnv*>(env)->functions[(size_t)i].name.c_str(); }
const char* fh_function_params(void* env, int i) { return static_cast<FakeEnv*>(env)->functions[(size_t)i].params.c_str(); }
long fh_frames_allocated(void* env) { return static_cast<FakeEnv*>(env)->frames_allocated; }

void* fh_source_create(int width, int height, int components, int sub_w, int sub_h, int bits,
                       int is_rgb, int is_planar, int num_frames, int parity_mode) {
    VideoInfo vi;
    vi.width = width; vi.height = height; vi.num_frames = num_frames;
    vi.stub_components = components; vi.stub_sub_w = sub_w; vi.stub_sub_h = sub_h; vi.stub_bits = bits;
    vi.stub_rgb = is_rgb != 0; vi.stub_planar = is_planar != 0;
    auto* h = new ClipHandle();
    h->source = new SourceClip(vi, parity_mode);
    h->clip = PClip(h->source);
    return h;
}

// A clip of `num_frames` frames that repeats `stored` distinct ones (frame n shows stored frame n % stored).
void* fh_source_create_looped(int width, int height, int components, int sub_w, int sub_h, int bits,
                              int is_rgb, int is_planar, int stored, int num_frames, int parity_mode) {
    VideoInfo vi;
    vi.width = width; vi.height = height; vi.num_frames = num_frames;
    vi.stub_components = components; vi.stub_sub_w = sub_w; vi.stub_sub_h = sub_h; vi.stub_bits = bits;
    vi.stub_rgb = is_rgb != 0; vi.stub_planar = is_planar != 0;
    auto* h = new ClipHandle();
    h->source = new SourceClip(vi, parity_mode, stored);
    h->clip = PClip(h->source);
    return h;
}

// Copy one plane of frame n into the source clip (src_pitch in bytes).
int fh_source_set_plane(void* clip, int n, int plane_index, const void* data, int src_pitch) {
    auto* h = static_cast<ClipHandle*>(clip);
    if (!h->source || n < 0 || n >= (int)h->source->frames.size()) return -1;
    VideoFrame* f = h->source->frames[(size_t)n].operator->();
    const int id = kPlaneIds[plane_index];
    BYTE* d = f->GetWritePtr(id);
    for (int y = 0; y < f->GetHeight(id); ++y)
        std::memcpy(d + (int64_t)y * f->GetPitch(id), static_cast<const BYTE*>(data) + (int64_t)y * src_pitch, (size_t)f->GetRowSize(id));
    return 0;
}

int fh_source_set_prop(void* clip, int n, const char* key, long long value) {
    auto* h = static_cast<ClipHandle*>(clip);
    if (!h->source || n < 0 || n >= (int)h->source->frames.size()) return -1;
    h->source->frames[(size_t)n]->stub_set_prop(key, value);
    return 0;
}

int fh_source_request_count(void* clip) { return (int)static_cast<ClipHandle*>(clip)->source->requests.size(); }
int fh_source_request_at(void* clip, int i) { return static_cast<ClipHandle*>(clip)->source->requests[(size_t)i]; }
void fh_source_clear_requests(void* clip) { static_cast<ClipHandle*>(clip)->source->requests.clear(); }
void fh_source_fail_at(void* clip, int n) { static_cast<ClipHandle*>(clip)->source->fail_at = n; }

void fh_clip_release(void* clip) { delete static_cast<ClipHandle*>(clip); }

// Call a registered function the way the script evaluator would: positional clip first, then
// named arguments; unspecified parameters are passed undefined. `values` are ints (bools as
// 0/1). Returns a new clip handle or NULL with err filled with the ThrowError message.
void* fh_invoke(void* env, const char* func, void* clip, int nargs, const char* const* names, const int* values,
                char* err, int errlen) {
    auto* e = static_cast<FakeEnv*>(env);
    const RegisteredFunction* f = nullptr;
    for (auto& g : e->functions) if (g.name == func) f = &g;
    if (!f) { set_err(err, errlen, "no such function"); return nullptr; }
    std::vector<AVSValue> argv(f->types.size());
    argv[0] = AVSValue(static_cast<ClipHandle*>(clip)->clip);
    for (int a = 0; a < nargs; ++a) {
        bool found = false;
        for (size_t k = 1; k < f->names.size(); ++k) {
            if (f->names[k] == names[a]) {
                argv[k] = f->types[k] == 'b' ? AVSValue(values[a] != 0) : AVSValue(values[a]);
                found = true;
            }
        }
        if (!found) {
            std::string m = std::string("Script error: ") + func + " does not have a named argument \"" + names[a] + "\"";
            set_err(err, errlen, m.c_str());
            return nullptr;
        }
    }
    try {
        AVSValue r = f->apply(AVSValue(argv.data(), (int)argv.size()), f->user_data, e);
        if (!r.IsClip()) { set_err(err, errlen, "function did not return a clip"); return nullptr; }
        auto* h = new ClipHandle();
        h->clip = r.AsClip();
        return h;
    } catch (const AvisynthError& x) {
        set_err(err, errlen, x.msg);
        return nullptr;
    }
}

int fh_clip_info(void* clip, int* out /* width,height,num_frames,components,sub_w,sub_h,bits */) {
    const VideoInfo& vi = static_cast<ClipHandle*>(clip)->clip->GetVideoInfo();
    out[0] = vi.width; out[1] = vi.height; out[2] = vi.num_frames; out[3] = vi.stub_components;
    out[4] = vi.stub_sub_w; out[5] = vi.stub_sub_h; out[6] = vi.stub_bits;
    return 0;
}

int fh_clip_cache_hints(void* clip, int hint, int range) {
    return static_cast<ClipHandle*>(clip)->clip->SetCacheHints(hint, range);
}

int fh_clip_parity(void* clip, int n) { return static_cast<ClipHandle*>(clip)->clip->GetParity(n) ? 1 : 0; }

void* fh_get_frame(void* env, void* clip, int n, char* err, int errlen) {
    auto* e = static_cast<FakeEnv*>(env);
    try {
        PVideoFrame f = static_cast<ClipHandle*>(clip)->clip->GetFrame(n, e);
        if (!f) { set_err(err, errlen, "null frame"); return nullptr; }
        auto* h = new FrameHandle();
        h->frame = f;
        return h;
    } catch (const AvisynthError& x) {
        set_err(err, errlen, x.msg);
        return nullptr;
    }
}

// out = pitch,row_size,height ; returns the read pointer
const void* fh_frame_plane(void* frame, int plane_index, int* out) {
    VideoFrame* f = static_cast<FrameHandle*>(frame)->frame.operator->();
    const int id = kPlaneIds[plane_index];
    out[0] = f->GetPitch(id); out[1] = f->GetRowSize(id); out[2] = f->GetHeight(id);
    return f->GetReadPtr(id);
}

int fh_frame_get_prop(void* frame, const char* key, long long* value) {
    int64_t v = 0;
    bool ok = static_cast<FrameHandle*>(frame)->frame->stub_get_prop(key, &v);
    *value = v;
    return ok ? 1 : 0;
}

void fh_frame_release(void* frame) { delete static_cast<FrameHandle*>(frame); }

}  // extern "C"
