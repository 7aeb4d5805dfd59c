NP=${NP:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $NP --master-addr 127.0.0.1"
for m in nvls p2p; do
KCNN_P2P_CHECK_MODE=$m timeout 120 $TR --master-port 29510 tools/p2p_check.py > gpurun_out/p2p_check_${m}_$NP.log 2>&1; echo "p2p_check $m rc=$?"
grep "p2p_check\|MISMATCH\|Error" gpurun_out/p2p_check_${m}_$NP.log | head -5
done
run() { tag=$1; shift; timeout 150 $TR --master-port 29513 bench.py --gpus $NP --steps 30 --warmup 5 --no-cpu --no-kernels "$@" > gpurun_out/bench_r1q_dp${NP}_$tag.json 2> gpurun_out/bench_r1q_dp${NP}_$tag.err; echo "bench $tag rc=$?"; }
run nvls --dp-reduce nvls
KCNN_P2P_CTAS=16 run nvls16 --dp-reduce nvls
run p2p --dp-reduce p2p
run nccl --dp-reduce nccl
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/bench_r1q_dp8*.json")+glob.glob("gpurun_out/bench_r1q_dp4*.json")):
    for l in open(f):
        if l.startswith("{"):
            d=json.loads(l); print(f, d["value"], d["ms_per_step"], d["objf_per_frame_last"], d["param_checksum"], d["e2e"]["value"], d["config"].get("dp_reduce"))
PY
grep -h "unavailable" gpurun_out/bench_r1q_dp${NP}*.err; true
