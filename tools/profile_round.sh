set -x
python bench.py > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_reference.json 2> gpurun_out/r02_bench_reference.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/r02_launches.csv python bench.py --quick --steps 1 --warmup 1 --no-graph --no-cpu-baseline > gpurun_out/r02_ncu_launch.log 2>&1
python tools/field_probe.py > gpurun_out/r02_probe.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:attn_bwd_tc -s 1 -c 1 -o gpurun_out/r02_prof_attn_bwd python tools/field_probe.py > gpurun_out/r02_ncu_a.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 8 -c 7 -o gpurun_out/r02_prof_gemms python tools/field_probe.py > gpurun_out/r02_ncu_g.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:attn_fwd_pp -s 1 -c 1 -o gpurun_out/r02_prof_attn_fwd python tools/field_probe.py > gpurun_out/r02_ncu_f.log 2>&1
python tools/sweeps.py --which 1,4,5 > gpurun_out/r02_sweeps.jsonl 2> gpurun_out/r02_sweeps.err
for w in c10b8 c10 distill infer sweep; do python bench.py --workload $w --no-cpu-baseline > gpurun_out/r02_bench_$w.json 2> gpurun_out/r02_bench_$w.err; done
python bench.py --workload c100 --precision fp32 --no-cpu-baseline --steps 3 > gpurun_out/r02_bench_c100_fp32.json 2> gpurun_out/r02_bench_c100_fp32.err
ls -la gpurun_out | tail -30
