python tools/step_ab.py 200
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
