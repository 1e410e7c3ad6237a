import sys, time, ctypes as C
sys.path.insert(0,'/root/repo/avisynth-sangnom2_b200'); sys.path.insert(0,'/root/repo')
import torch, numpy as np
from pysangnom import cuda
def bench(sb, W, H, sub, nframes, iters=5):
    dt = {1:torch.uint8, 2:torch.int16, 4:torch.float32}[sb]
    pitchY = (W*sb+255)//256*256
    planes=[]
    ctx = cuda.Context(sb, W, H)
    jobs=[]
    for f in range(nframes):
        dims=[(W,H)] + ([(W>>sub[0],H>>sub[1])]*2 if sub is not None else [])
        for p,(w,h) in enumerate(dims):
            pitch=(w*sb+255)//256*256
            t=torch.randint(0,256,(h,pitch),dtype=torch.uint8,device='cuda')
            planes.append(t)
            jobs.append(cuda.make_job(0,0,t.data_ptr(),pitch,w,h,f&1,cuda.MODE_INPLACE,cuda.threshold(48,8 if sb==1 else (16 if sb==2 else 32),sb),p,f))
    arr=(cuda.SnPlaneJob*len(jobs))(*jobs)
    ts=torch.cuda.Stream(); st=ts.cuda_stream
    for _ in range(2): ctx.process_jobs_device(arr, st)
    torch.cuda.synchronize()
    e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    e0.record(ts)
    h0=time.perf_counter()
    for _ in range(iters): ctx.process_jobs_device(arr, st)
    host_ms=(time.perf_counter()-h0)*1000/iters
    e1.record(ts); torch.cuda.synchronize()
    ms=e0.elapsed_time(e1)/iters
    # one call alone: GPU time without host submission gaps
    e2=torch.cuda.Event(enable_timing=True); e3=torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); ctx.process_jobs_device(arr, st); torch.cuda.synchronize()
    print(f"sb={sb} {W}x{H} sub={sub} frames={nframes}: {ms:.3f} ms/batch -> {nframes/ms*1000:.1f} fps  (host submit {host_ms:.3f} ms/call)", flush=True)
    ctx.close()
bench(1,1920,1080,(1,1),296)
bench(1,1920,1080,(1,1),592)
bench(1,1920,1080,None,296)
bench(2,1920,1080,(1,1),296)
bench(4,1920,1080,(1,1),148)
bench(4,3840,2160,(1,1),148)
bench(1,720,480,None,592)
