#!/bin/bash
# Round-2 evidence run (one B200): tests, bench lines of every BASELINE config, kernel microbenchmarks, ncu launch list of the
# bench step and ncu --set full captures per kernel family.  Everything lands in gpurun_out/r2p_*; summaries are copied to profiles/.
set -x
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > $O/r2p_tests.log 2>&1; echo "rc=$?" >> $O/r2p_tests.log
timeout 300 python __graft_entry__.py smoke > $O/r2p_smoke.log 2>&1
timeout 600 python bench.py --steps 20 --warmup 3 > $O/r2p_bench.json 2> $O/r2p_bench.err
for cfg in "--deg inpaint_random --chains 1" "--deg deblur_aniso" "--deg phase_retrieval" "--workload latent --deg inpaint_random" "--deg sr16" "--deg hdr"; do
  timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-fp32-line $cfg >> $O/r2p_config_lines.jsonl 2>> $O/r2p_config_lines.err
done
timeout 600 python tools/kernel_microbench.py --reps 10 > $O/r2p_mb_cold.jsonl 2> $O/r2p_mb_cold.err
timeout 300 python tools/kernel_microbench.py --only gn --small --warm --graph-calls 20 --reps 5 > $O/r2p_mb_small.jsonl 2>&1
# ncu: launch list of the bench step (cold-cache, serialised: compare shares)
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-in-situ --no-fp32-line --cuda-profiler-range > $O/r2p_plain_bench.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $O/r2p_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-in-situ --no-fp32-line --cuda-profiler-range > $O/r2p_ncu_bench.log 2>&1
# ncu --set full per family
python tools/kernel_microbench.py --only gn --shape 8,128,256 --reps 1 --resident 0 > /dev/null 2>&1 && \
ncu --set full --clock-control none -k regex:gn_nhwc -c 12 -o /tmp/r2p_gn python tools/kernel_microbench.py --only gn --shape 8,128,256 --reps 1 --resident 0 > $O/r2p_ncu_gn.log 2>&1
python tools/kernel_microbench.py --only leapfrog --reps 1 > /dev/null 2>&1 && \
ncu --set full --clock-control none -k regex:leapfrog_kernel -s 9 -c 3 -o /tmp/r2p_leapfrog python tools/kernel_microbench.py --only leapfrog --reps 1 > $O/r2p_ncu_lf.log 2>&1
python tools/kernel_microbench.py --only ddim --reps 1 > /dev/null 2>&1 && \
ncu --set full --clock-control none -k regex:ddim_ -s 6 -c 2 -o /tmp/r2p_ddim python tools/kernel_microbench.py --only ddim --reps 1 > $O/r2p_ncu_ddim.log 2>&1
python tools/kernel_microbench.py --only op --reps 1 > /dev/null 2>&1 && \
ncu --set full --clock-control none -k regex:"loss_grad|tc_gemm|phase_" -c 30 -o /tmp/r2p_ops python tools/kernel_microbench.py --only op --reps 1 > $O/r2p_ncu_ops.log 2>&1
python tools/kernel_microbench.py --only attn --reps 1 > /dev/null 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:attn_fwd_kernel -c 1 -o /tmp/r2p_attn python tools/kernel_microbench.py --only attn --reps 1 > $O/r2p_ncu_attn.log 2>&1
ncu -i /tmp/r2p_attn.ncu-rep --page source --csv > $O/r2p_attn_source.csv 2>/dev/null
# the .ncu-rep files stay on the box (gpurun_out is capped at 64 MiB): only the raw-page CSVs travel
for r in gn leapfrog ddim ops attn; do ncu -i /tmp/r2p_$r.ncu-rep --page raw --csv > $O/r2p_${r}_raw.csv 2>/dev/null; done
du -sh $O
ls -la $O | tail -30
