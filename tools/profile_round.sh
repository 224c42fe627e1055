set -x
cd $GRAFT_REPO_ROOT
# 1. launch list of two steps (graph nodes) with DRAM bytes
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r01_launches_raw.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-breakdown > gpurun_out/r01_launches_bench.log 2>&1
tail -c 300 gpurun_out/r01_launches_bench.log
wc -l gpurun_out/r01_launches_raw.csv
# 2. full captures of the hot kernels
ncu --set full --clock-control none --import-source on -k regex:"conv_tc" -s 3 -c 3 -o gpurun_out/r01_prof_conv_final -f python tools/run_ops.py conv > gpurun_out/r01_ncu_conv.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"fact_mma" -s 2 -c 2 -o gpurun_out/r01_prof_fact_final -f python tools/run_ops.py fact > gpurun_out/r01_ncu_fact.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"attention_f16" -s 1 -c 1 -o gpurun_out/r01_prof_attn_final -f python tools/run_ops.py attn > gpurun_out/r01_ncu_attn.log 2>&1
ls -la gpurun_out/*.ncu-rep
# 3. final bench lines
python bench.py --steps 20 --warmup 3 --dump-breakdown gpurun_out/r01_breakdown_final.csv > gpurun_out/r01_bench_final.json 2> gpurun_out/r01_bench_final.err
head -c 600 gpurun_out/r01_bench_final.json
