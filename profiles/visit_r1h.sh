timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
bash profiles/ab.sh r1h "-" "FSMG_STRIP_STREAM=2" "FSMG_STRIP_STREAM=0" "FSMG_STRIP_STREAM=0 FSMG_STRIP_LOOP=32" "-"
