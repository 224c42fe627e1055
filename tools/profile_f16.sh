set -x
cd $GRAFT_REPO_ROOT
# 1. launch list of the timed steps (graph nodes) with DRAM bytes
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r01_launches_f16_raw.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-breakdown --no-vae > gpurun_out/r01_launches_f16_bench.log 2>&1
tail -c 300 gpurun_out/r01_launches_f16_bench.log
wc -l gpurun_out/r01_launches_f16_raw.csv
# 2. full capture of the fp16-operand conv kernel (3 layers)
ncu --set full --clock-control none --import-source on -k regex:"conv_tc" -s 3 -c 3 -o gpurun_out/r01_prof_conv_f16 -f python tools/run_ops.py conv16 > gpurun_out/r01_ncu_conv_f16.log 2>&1
ls -la gpurun_out/*.ncu-rep
