bash profiles/ab.sh r1g "-" "FSMG_STRIP_LOOP=8" "FSMG_STRIP_LOOP=32" "FSMG_STRIP_LOOP=0" 
