# ncu evidence for the training step (run under gpurun): launch list of two steps + full captures of the new tensor-core kernels
set -x
python tools/train_one.py > gpurun_out/train_one_plain.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 12000 --csv --log-file gpurun_out/r02_train_launches.csv python tools/train_one.py > gpurun_out/train_one_ncu.log 2>&1
N="ncu --set full --import-source on --clock-control none --kernel-name-base demangled"
$N -k regex:"attention_bwd_kernel<true>|attention_bwd_kernel<\(bool\)1>" -s 15 -c 1 -o gpurun_out/k_attn_bwd_dkv -f python tools/train_one.py > /dev/null 2>&1
$N -k regex:"attention_bwd_kernel<false>|attention_bwd_kernel<\(bool\)0>" -s 15 -c 1 -o gpurun_out/k_attn_bwd_dq -f python tools/train_one.py > /dev/null 2>&1
$N -k regex:gemm_tn_kernel -s 100 -c 1 -o gpurun_out/k_gemm_tn -f python tools/train_one.py > /dev/null 2>&1
ls -la gpurun_out/k_attn_bwd_*.ncu-rep gpurun_out/k_gemm_tn.ncu-rep
