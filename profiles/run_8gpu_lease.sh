#!/bin/bash
# Round 2: everything that needs all GPUs of one box, in one lease (VERDICT r1 "next round" items 1 and 5).
# Writes JSONL files under gpurun_out/ (copied to profiles/r2_*.jsonl afterwards).
NG=$(nvidia-smi -L | wc -l)
O=gpurun_out
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
port=29600
# (b) headline strong scaling: 10^8 points in total, ceil(n/G) per rank, + timed all-gather of the 5.6 GB result (C2)
: > $O/r2_strong_scaling.jsonl
for G in 1 2 4 8; do
  [ $G -gt $NG ] && break
  port=$((port+1))
  if [ $G -eq 1 ]; then CUDA_VISIBLE_DEVICES=0 python bench.py --gpus 1 --steps 5 --warmup 3 --scaling strong --no-cpu-baseline >> $O/r2_strong_scaling.jsonl 2>> $O/r2_lease.err
  else $TR --nproc-per-node $G --master-port $port bench.py --gpus $G --steps 5 --warmup 3 --scaling strong --gather >> $O/r2_strong_scaling.jsonl 2>> $O/r2_lease.err; fi
done
# (a) cfg 5 as stated: 10^7 points x (K x 256) across all GPUs, device-resident shards + gather
port=$((port+1))
$TR --nproc-per-node $NG --master-port $port profiles/run_cfg5_torchrun.py 10000000 --gather > $O/r2_cfg5.jsonl 2>> $O/r2_lease.err
# the same through the one-process C-ABI entries: device-resident shards with the in-library NCCL gather, then host buffers
python profiles/run_sharded_device.py cfg5 10000000 >> $O/r2_cfg5.jsonl 2>> $O/r2_lease.err
python profiles/run_sharded_device.py cfg4 100000000 > $O/r2_sharded_device_cfg4.jsonl 2>> $O/r2_lease.err
python profiles/run_block_sharded.py 1250000 1 >> $O/r2_cfg5.jsonl 2>> $O/r2_lease.err
# (5) host<->device aggregate with 1, 2, 4, 8 GPUs copying at once
profiles/pcie_multi.sh 2 > $O/r2_pcie_concurrent.jsonl 2>> $O/r2_lease.err
# e2e at all GPUs with two chunk sizes of the host pipeline (weak, 2*10^7 points per GPU)
: > $O/r2_e2e_chunks.jsonl
for CH in 2097152 8388608; do
  port=$((port+1))
  SK_HOST_CHUNK=$CH $TR --nproc-per-node $NG --master-port $port bench.py --gpus $NG --steps 3 --warmup 3 --points 20000000 >> $O/r2_e2e_chunks.jsonl 2>> $O/r2_lease.err
done
tail -c 1500 $O/r2_lease.err
wc -c $O/r2_*.jsonl
