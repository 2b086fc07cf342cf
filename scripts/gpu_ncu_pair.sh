# ncu --set full capture of the CTA-pair convolution kernels:  gpurun --timeout 900 -- 'bash scripts/gpu_ncu_pair.sh'
set -x
mkdir -p gpurun_out
timeout 120 python scripts/ncu_pair_layers.py > gpurun_out/pair_layers_plain.log 2>&1 || exit 1
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:conv_ -o gpurun_out/r2_full_pair -f python scripts/ncu_pair_layers.py > gpurun_out/ncu_full_pair.log 2>&1; tail -2 gpurun_out/ncu_full_pair.log | cut -c1-200
ls -la gpurun_out/*.ncu-rep
