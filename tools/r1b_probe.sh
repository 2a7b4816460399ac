# session-2 probe: GPU tests + in-kernel timelines of the headline shapes
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for s in "2048 4096 4096" "2048 11008 4096" "2048 4096 11008"; do
  for cfg in 1 0; do python tools/timeline.py $s $cfg 0; done
done
for s in "2048 4096 4096" "2048 11008 4096" "2048 4096 11008"; do
  for cfg in 1 0; do python tools/prof_gemm.py $s $cfg 20 0; done
  PQ_STAGED=1 python tools/prof_gemm.py $s -1 20 0
done
