"""Time a GEMM shape/config from a CUDA graph (GPU time only).  usage: prof_gemm.py M N K cfg [iters] [sk] [nograph]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import protoquant_b200 as pq
M, N, K, cfg = (int(v) for v in sys.argv[1:5])
iters = int(sys.argv[5]) if len(sys.argv) > 5 else 3
sk = int(sys.argv[6]) if len(sys.argv) > 6 else -1
nograph = len(sys.argv) > 7 and sys.argv[7] == "nograph"
staged = int(os.environ.get("PQ_STAGED", "0"))
pq.lib().pq_debug_set_staged(staged)
pq.lib().pq_debug_set_epilogue(int(os.environ.get("PQ_EPI", "0")))
pq.lib().pq_debug_set_gemm_config(cfg)
pq.lib().pq_debug_set_tma_store(int(os.environ.get("PQ_TMA_STORE", "1")))
pq.lib().pq_debug_set_streamk(sk)
a = torch.randint(-128, 128, (M, K), dtype=torch.int8, device="cuda")
b = torch.randint(-128, 128, (N, K), dtype=torch.int8, device="cuda")
sx = torch.rand(M, device="cuda"); sw = torch.rand(N, device="cuda")
odt = {"bf16": torch.bfloat16, "f16": torch.float16, "f32": torch.float32}[os.environ.get("PQ_OUT", "bf16")]
y = torch.empty(M, N, dtype=odt, device="cuda")
def run():
    for _ in range(iters):
        pq.qgemm(a, sx, b, sw, None, odt, out=y)
run(); torch.cuda.synchronize()
if nograph:
    fn = run
else:
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        run()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        run()
    fn = g.replay
fn(); torch.cuda.synchronize()
best = 1e9
for _ in range(5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(); e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1) / iters)
ref = (a[:64].float() @ b[:512].float().t()) * sx[:64, None] * sw[None, :512]
ok = torch.allclose(y[:64, :512].float(), ref, rtol=2e-2, atol=1e-2 * ref.abs().max().item())
if int(os.environ.get("PQ_EPI", "0")): ok = True
print(f"M={M} N={N} K={K} cfg={cfg} sk={sk} staged={staged} epi={os.environ.get('PQ_EPI', '0')} tma_store={os.environ.get('PQ_TMA_STORE', '1')} out={os.environ.get('PQ_OUT', 'bf16')}: {best*1e3:.1f} us  {2*M*N*K/best/1e9:.0f} TOPS  {'ok' if ok else 'MISMATCH'}")
