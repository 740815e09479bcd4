set -x
mkdir -p gpurun_out/c20
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attention_prefill_tc -s 3 -c 1 -o gpurun_out/c20/ncu_attn_tc python tools/ncu_target_attn.py > gpurun_out/c20/ncu.log 2>&1
tail -n 3 gpurun_out/c20/ncu.log
