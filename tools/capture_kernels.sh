set -x
B="python bench.py --steps 2 --warmup 3 --cpu-sample 0 --e2e-steps 1 --no-extras --no-gpu-baseline"
N="ncu --set full --import-source on --clock-control none --kernel-name-base demangled"
$N -k regex:"tc2_kernel.*256.*1, s3od::EpiConv" -s 40 -c 1 -o gpurun_out/k_pairconv -f $B > /dev/null 2>&1
$N -k regex:"tc2_kernel.*128.*1, s3od::EpiConv" -s 3 -c 1 -o gpurun_out/k_mhc1 -f $B > /dev/null 2>&1
$N -k regex:EpiQKV -s 40 -c 1 -o gpurun_out/k_qkv -f $B > /dev/null 2>&1
$N -k regex:EpiResidual -s 81 -c 1 -o gpurun_out/k_residual -f $B > /dev/null 2>&1
$N -k regex:layernorm_kernel -s 80 -c 1 -o gpurun_out/k_ln -f $B > /dev/null 2>&1
$N -k regex:convt_rows -s 6 -c 1 -o gpurun_out/k_convt -f $B > /dev/null 2>&1
$N -k regex:"conv_rows_kernel.*96" -s 3 -c 1 -o gpurun_out/k_heads -f $B > /dev/null 2>&1
$N -k regex:"conv_rows_kernel.*64" -s 3 -c 1 -o gpurun_out/k_rows64 -f $B > /dev/null 2>&1
ls -la gpurun_out/k_*.ncu-rep
