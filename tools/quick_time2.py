import sys, time
sys.path.insert(0,'/root/repo/avisynth-sangnom2_b200'); sys.path.insert(0,'/root/repo')
import torch
from pysangnom import cuda
def bench(sb, W, H, sub, nframes, mode, pitch_align=256, iters=5):
    ctx = cuda.Context(sb, W, H)
    jobs=[]; keep=[]
    for f in range(nframes):
        dims=[(W,H)] + ([(W>>sub[0],H>>sub[1])]*2 if sub is not None else [])
        for p,(w,h) in enumerate(dims):
            pitch=(w*sb+pitch_align-1)//pitch_align*pitch_align
            t=torch.randint(0,256,(h,pitch),dtype=torch.uint8,device='cuda')
            keep.append(t)
            thr=cuda.threshold(48,8 if sb==1 else (16 if sb==2 else 32),sb)
            if mode=='inplace':
                jobs.append(cuda.make_job(0,0,t.data_ptr(),pitch,w,h,f&1,cuda.MODE_INPLACE,thr,p,f))
            else:
                d=torch.empty_like(t); keep.append(d)
                jobs.append(cuda.make_job(t.data_ptr(),pitch,d.data_ptr(),pitch,w,h,f&1,cuda.MODE_FIELD,thr,p,f))
    arr=(cuda.SnPlaneJob*len(jobs))(*jobs)
    ts=torch.cuda.Stream(); st=ts.cuda_stream
    for _ in range(2): ctx.process_jobs_device(arr, st)
    torch.cuda.synchronize()
    e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    e0.record(ts)
    for _ in range(iters): ctx.process_jobs_device(arr, st)
    e1.record(ts); torch.cuda.synchronize()
    ms=e0.elapsed_time(e1)/iters
    print(f"sb={sb} {W}x{H} sub={sub} frames={nframes} mode={mode} pitch_align={pitch_align}: {ms:.3f} ms/batch -> {nframes/ms*1000:.1f} fps", flush=True)
    ctx.close()
for mode in ('inplace','field'):
    for n in (128, 296, 592):
        bench(1,1920,1080,(1,1),n,mode)
bench(1,1920,1080,(1,1),128,'field',16)
bench(1,1920,1080,(1,1),128,'inplace',16)
