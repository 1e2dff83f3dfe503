timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
bash profiles/ab.sh r1d "-" "FSMG_STRIP_OVERLAP=0" "FSMG_STRIP_PER_SM=2" "FSMG_STRIP_PER_SM=6" "FSMG_STRIP_OVERLAP=0 FSMG_LSTM_ROT=2" "FSMG_STRIP_OVERLAP=0 FSMG_LSTM_ROT=0" "FSMG_STRIP_OVERLAP=0 FSMG_LSTM_PAIR=0"
FSMG_COOP=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r1d_launches.csv python profiles/profile_step.py 2 > gpurun_out/r1d_launches.log 2>&1
tail -3 gpurun_out/r1d_launches.log
