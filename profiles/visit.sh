bash profiles/ab.sh r2i "-" "FSMG_ASTAT=1" "-" "FSMG_ASTAT=1"
