# how profiles/r2_bench_n1_*.json, r2_launches.csv and r2_kgen_raw.csv were produced: gpurun -- 'bash tools/run_profile.sh'
mkdir -p gpurun_out
python bench.py > gpurun_out/r2f_n1_1m_x16.out 2> gpurun_out/r2f_n1_1m_x16.err; echo rc=$?; tail -1 gpurun_out/r2f_n1_1m_x16.out > gpurun_out/r2f_n1_1m_x16.json
for w in 200k_x8 20k_x2 64ch_60k_x4; do python bench.py --workload $w > gpurun_out/r2f_n1_$w.out 2> gpurun_out/r2f_n1_$w.err; echo rc=$?; tail -1 gpurun_out/r2f_n1_$w.out > gpurun_out/r2f_n1_$w.json; done
python bench.py --mode launch --no-extras > gpurun_out/r2f_n1_launch.out 2>&1; tail -1 gpurun_out/r2f_n1_launch.out > gpurun_out/r2f_n1_launch.json
python bench.py --impl reference > gpurun_out/r2f_ref_n1.out 2>&1; tail -1 gpurun_out/r2f_ref_n1.out > gpurun_out/r2f_ref_n1.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2f_launches.csv python bench.py --no-cpu-baseline > gpurun_out/r2f_ncu_l.log 2>&1; echo ncu-list rc=$?
ncu --set full --clock-control none -k regex:k_generations -c 3 -o gpurun_out/r2f_kgen -f python bench.py --steps 20 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2f_ncu_f.log 2>&1; echo ncu-full rc=$?
ncu -i gpurun_out/r2f_kgen.ncu-rep --page raw --csv > gpurun_out/r2f_kgen_raw.csv; ls -la gpurun_out/r2f_kgen.ncu-rep; rm -f gpurun_out/r2f_kgen.ncu-rep; du -sh gpurun_out
