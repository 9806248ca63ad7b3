set -x
python bench.py --steps 10 --torch-profile gpurun_out/tp_final.txt > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-in-situ --cuda-profiler-range > gpurun_out/plain_bench_r1.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_r1c.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-in-situ --cuda-profiler-range > gpurun_out/ncu_bench.log 2>&1
python tools/kernel_microbench.py --reps 10 > gpurun_out/mb_cold.jsonl 2> gpurun_out/mb_cold.err
python tools/kernel_microbench.py --only gn --small --warm --graph-calls 20 --reps 5 > gpurun_out/mb_small_final.jsonl 2>&1
python tools/kernel_microbench.py --only gn --shape 8,512,16 --reps 1 > /dev/null 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:gn_nhwc_slice -c 8 -o gpurun_out/gn_solo_r1 python tools/kernel_microbench.py --only gn --shape 8,512,16 --reps 1 > gpurun_out/ncu_solo.log 2>&1
ncu -i gpurun_out/gn_solo_r1.ncu-rep --page raw --csv > gpurun_out/gn_solo_r1_raw.csv 2>/dev/null
ls -la gpurun_out | tail -20
