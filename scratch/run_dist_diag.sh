set -u
for k in "" a b c; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 500)) tests/_dist_worker.py sync $k 2>/dev/null | grep RESULT
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 500)) tests/_dist_worker.py ddp 2>/dev/null | grep RESULT
