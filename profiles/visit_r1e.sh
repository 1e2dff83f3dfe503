bash profiles/ab.sh r1e "-" "FSMG_CHUNK_ROWS=18432" "FSMG_CHUNK_ROWS=9216" "FSMG_CHUNK_ROWS=18944" "FSMG_CHUNK_ROWS=36864" "FSMG_STRIP_ROWS=16" "FSMG_STRIP_ROWS=64"
FSMG_TRACE=1 timeout 120 python profiles/profile_step.py 1 > gpurun_out/r1e_trace.log 2>&1
tail -30 gpurun_out/r1e_trace.log
