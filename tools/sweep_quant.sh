python tools/prof_quant.py 2048 4096 bf16 0 0 128 4 64 8 256 2 128 3 64 6
python tools/prof_quant.py 2048 11008 bf16 0 0 512 4 256 6 512 3 256 8 1024 2
python tools/prof_quant.py 65536 11008 bf16 0 0 512 4 256 6 512 3 256 8
python tools/prof_quant.py 32768 28672 bf16 0 0 1024 4 512 8 1024 6
python tools/prof_quant.py 131072 4096 bf16 0 0 128 4 64 8 256 2
python tools/prof_quant.py 4096 768 bf16 0 0 32 3 32 4 64 2
python tools/prof_quant.py 65536 11008 f32 0 0 1024 4 512 6 512 8 1024 3
python tools/prof_quant.py 131072 4096 f32 0 0
python tools/prof_quant.py 131072 4096 f16 0 0
for st in 0 1; do PQS=$st python - <<'PY'
import os, sys
sys.path.insert(0, os.getcwd())
import torch, protoquant_b200 as pq
st = int(os.environ["PQS"]); pq.lib().pq_debug_set_staged(st)
for (M,N,K) in ((2048,4096,4096),(2048,11008,4096),(2048,4096,11008),(8192,8192,8192)):
    a = torch.randint(-128,128,(M,K),dtype=torch.int8,device="cuda"); b = torch.randint(-128,128,(N,K),dtype=torch.int8,device="cuda")
    sx = torch.rand(M,device="cuda"); sw = torch.rand(N,device="cuda"); y = torch.empty(M,N,dtype=torch.bfloat16,device="cuda")
    def run():
        for _ in range(20): pq.qgemm(a,sx,b,sw,None,torch.bfloat16,out=y)
    run(); torch.cuda.synchronize()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s): run()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g): run()
    g.replay(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(5):
        e0,e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize(); best=min(best,e0.elapsed_time(e1)/20)
    print(f"staged={st} {M}x{N}x{K}: {best*1e3:.1f} us {2*M*N*K/best/1e9:.0f} TOPS")
PY
done
