timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
bash profiles/ab.sh r1t "-" "FSMG_DWS_T=0" "-" "FSMG_DWS_T=0"
