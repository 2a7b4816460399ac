for s in "8192 8192 8192" "4096 4096 16384" "8192 4096 8192"; do
  for cfg in 1 16 0; do python tools/clock_probe.py $s $cfg 2.5; done
done
