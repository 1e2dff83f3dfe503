timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
bash profiles/ab.sh r2k "-" "-"
python profiles/eval_phases.py 2>&1 | tail -2
