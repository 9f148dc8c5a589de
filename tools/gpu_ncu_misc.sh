cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
python tools/profile_misc.py > gpurun_out/r2v_misc_plain.log 2>&1 || { tail -5 gpurun_out/r2v_misc_plain.log; exit 1; }
ncu --clock-control none --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum -c 200 --csv --log-file gpurun_out/r2_launches_misc.csv python tools/profile_misc.py > /dev/null 2>&1
ncu --clock-control none --set full --import-source on --kernel-id '::regex:.*:2' -o gpurun_out/prof_r2_misc -f python tools/profile_misc.py > gpurun_out/r2v_misc_ncu.log 2>&1
ls -la gpurun_out/prof_r2_misc.ncu-rep; tail -2 gpurun_out/r2v_misc_ncu.log
