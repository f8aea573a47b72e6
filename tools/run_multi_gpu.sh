#!/bin/bash
# measurements that need the 8 GPUs of one box (run under `gpurun --gpus 8`): the box's concurrent pinned-copy ceiling at
# N = 2 / 4 / 8, the default bench at N = 8, BASELINE.json configs[3] (vox11, 300 frames = 38 per GPU) and configs[4]
# (8 streams r1-r5).  Every JSON line lands in gpurun_out/.
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
for n in 2 4 8; do
  $TR --nproc-per-node $n --master-port 2950$n tools/pcie_peak.py 2>/dev/null | tail -1 > gpurun_out/pcie_peak_n$n.json
done
python tools/pcie_peak.py 2>/dev/null | tail -1 > gpurun_out/pcie_peak_n1_8gpubox.json
cat gpurun_out/pcie_peak_n*.json
timeout 900 $TR --nproc-per-node 8 --master-port 29611 bench.py --gpus 8 --steps 6 --warmup 3 --quick > gpurun_out/bench_n8_quick.json 2> gpurun_out/bench_n8_quick.err
echo "bench n8 rc=$?"; cut -c1-300 gpurun_out/bench_n8_quick.json
timeout 900 $TR --nproc-per-node 8 --master-port 29612 bench.py --gpus 8 --workload streams --frames 32 --steps 4 > gpurun_out/streams_n8.json 2> gpurun_out/streams_n8.err
echo "streams n8 rc=$?"; cut -c1-300 gpurun_out/streams_n8.json
timeout 1200 $TR --nproc-per-node 8 --master-port 29613 bench.py --gpus 8 --workload vox11 --frames 38 --distinct-frames 6 --steps 3 --quick --no-cpu-baseline > gpurun_out/vox11_n8.json 2> gpurun_out/vox11_n8.err
echo "vox11 n8 rc=$?"; cut -c1-300 gpurun_out/vox11_n8.json
tail -3 gpurun_out/*_n8*.err
