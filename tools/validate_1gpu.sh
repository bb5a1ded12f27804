#!/bin/bash
# Round-end validation on ONE B200 (run under gpurun): bench of cfg2 (with the `also` variants) and cfg4, the ncu launch list and
# one ncu --set full pass over the library's kernels.  Outputs land in gpurun_out/ (copy what should be judged into profiles/).
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_v4.json 2> gpurun_out/bench.err; echo "bench rc $?" > gpurun_out/rc.log
timeout 300 python bench.py --steps 5 --warmup 3 --experts 8 --topk 2 --img 384 --no-also --no-cpu-baseline > gpurun_out/r2_bench_cfg4_b256_v2.json 2> gpurun_out/bench4.err; echo "cfg4 rc $?" >> gpurun_out/rc.log
python bench.py --steps 2 --warmup 3 --no-also --no-cpu-baseline --no-graph --sustain-s 0 > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_ncu_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-also --no-cpu-baseline --no-graph --sustain-s 0 > gpurun_out/ncu1.log 2>&1; echo "ncu1 rc $?" >> gpurun_out/rc.log
python tools/profile_step.py > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none -k regex:"b2b|cm_|gemm_|bwd_z|rank1|dispatch|router|expert_reduce|global_mean|pack_params" -c 45 \
    -o /tmp/r2_full_step python tools/profile_step.py > gpurun_out/ncu2.log 2>&1; echo "ncu2 rc $?" >> gpurun_out/rc.log
# the report itself is too large to bring back (gpurun_out is capped at 64 MiB): keep the per-launch summary
ncu -i /tmp/r2_full_step.ncu-rep --page raw --csv > /tmp/r2_full_raw.csv 2>/dev/null && python tools/ncu_summary.py /tmp/r2_full_raw.csv > gpurun_out/r2_ncu_full_step.csv
du -sh gpurun_out; cat gpurun_out/rc.log
