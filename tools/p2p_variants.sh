cd /root/repo
for fl in "-DS3OD_P2P_TEST=0" "-DS3OD_P2P_TEST=1" "-DS3OD_P2P_TEST=2"; do
  touch s3od_b200/csrc/train.cu
  S3OD_NVCC_FLAGS="$fl" python -m s3od_b200.build > /dev/null 2>&1 || { echo "build failed $fl"; continue; }
  echo "== $fl"
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/p2p_probe.py 2>&1 | tail -1
done
touch s3od_b200/csrc/train.cu; python -m s3od_b200.build > /dev/null 2>&1
python -m pytest tests/test_gpu_training.py -q -x -k "emulated" 2>&1 | tail -12
