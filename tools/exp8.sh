cd $GRAFT_REPO_ROOT
python tools/pcie_peak.py --gpus 1,2,4,8 --out gpurun_out/r02_pcie_ceiling.json 2>gpurun_out/pcie.err | tail -4
for i in 1 2; do timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $((29500+i)) bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2_n8_$i.json 2> gpurun_out/r2_n8_$i.err; echo run $i rc=$?; done
tail -c 300 gpurun_out/r2_n8_1.err
