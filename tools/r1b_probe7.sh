for s in "2048 4096 4096" "2048 11008 4096"; do
  for cfg in 0 1 11 8 16; do python tools/clock_probe.py $s $cfg 2.5; done
done
python tools/clock_probe.py 2048 4096 11008 1 2.5
python tools/clock_probe.py 2048 4096 11008 0 2.5
python tools/clock_probe.py 2048 4096 11008 16 2.5
