set -x
mkdir -p gpurun_out/r02
(nproc; nvidia-smi topo -m; lscpu | grep -E "Model name|Socket|NUMA|^CPU\(s\)"; free -g | head -2) > gpurun_out/r02/host_info_8gpu.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 300 $TR --nproc-per-node 8 --master-port 29511 tools/host_copy_bw.py > gpurun_out/r02/host_copy_bw_8gpu.json 2> gpurun_out/r02/host_copy_bw_8gpu.err
timeout 300 $TR --nproc-per-node 4 --master-port 29512 tools/host_copy_bw.py > gpurun_out/r02/host_copy_bw_4gpu.json 2> gpurun_out/r02/host_copy_bw_4gpu.err
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 --no-configs --no-sweep --no-cpu > gpurun_out/r02/bench8_n1.json 2> gpurun_out/r02/bench8_n1.err
timeout 400 $TR --nproc-per-node 8 --master-port 29513 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r02/bench8_n8.json 2> gpurun_out/r02/bench8_n8.err
FLOCK_PIN_CORES=0 timeout 400 $TR --nproc-per-node 8 --master-port 29514 bench.py --gpus 8 --steps 20 --warmup 5 --no-configs --no-sweep > gpurun_out/r02/bench8_n8_nopin.json 2> gpurun_out/r02/bench8_n8_nopin.err
FLOCK_E2E_DEPTH=2 timeout 400 $TR --nproc-per-node 8 --master-port 29515 bench.py --gpus 8 --steps 20 --warmup 5 --no-configs --no-sweep > gpurun_out/r02/bench8_n8_depth2.json 2> gpurun_out/r02/bench8_n8_depth2.err
timeout 400 $TR --nproc-per-node 8 --master-port 29516 bench.py --gpus 8 --steps 20 --warmup 5 --workload cfg3 --no-configs --no-sweep > gpurun_out/r02/bench8_cfg3_n8.json 2> gpurun_out/r02/bench8_cfg3_n8.err
tail -2 gpurun_out/r02/*8*.err
