timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
bash profiles/ab.sh r2e "-" "-"
FSMG_COOP=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r2e_launches.csv python profiles/profile_step.py 1 > gpurun_out/r2e_launches.log 2>&1
