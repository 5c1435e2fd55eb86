timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/check_sharded.py 2>&1 | grep -v "^$\|Warning\|warn" | tail -12
for mode in peer nccl; do
echo "== exchange=$mode"
ERIRT_EXCHANGE=$mode timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --short --steps 50 --warmup 5 2>&1 | grep -o '"ms_per_step": [0-9.]*\|"value": [0-9.]*\|rror.*' | head -4
done
