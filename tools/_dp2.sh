NP=${NP:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $NP --master-addr 127.0.0.1"
for m in nvls p2p; do
KCNN_P2P_CHECK_MODE=$m timeout 200 $TR --master-port 29510 tools/p2p_check.py > gpurun_out/p2p_check_${m}_$NP.log 2>&1; echo "p2p_check $m rc=$?"
grep "p2p_check\|MISMATCH\|Error" gpurun_out/p2p_check_${m}_$NP.log | head -5
done
KCNN_DP_CHECK_REDUCE=nvls timeout 200 $TR --master-port 29511 tools/dp_check.py > gpurun_out/dp_check_nvls_$NP.log 2>&1; echo "dp_check nvls rc=$?"
grep "dp_check" gpurun_out/dp_check_nvls_$NP.log
for m in nvls p2p nccl; do
timeout 200 $TR --master-port 29513 bench.py --gpus $NP --steps 30 --warmup 5 --dp-reduce $m --no-cpu > gpurun_out/bench_r1q_dp${NP}_$m.json 2> gpurun_out/bench_r1q_dp${NP}_$m.err; echo "bench $m rc=$?"
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/bench_r1q_dp*.json")):
    for l in open(f):
        if l.startswith("{"):
            d=json.loads(l); print(f, d["value"], d["ms_per_step"], d["objf_per_frame_last"], d["param_checksum"], d["e2e"]["value"], d["config"].get("dp_reduce"))
PY
grep -h "unavailable" gpurun_out/bench_r1q_dp*.err
