set -x
mkdir -p gpurun_out/c9
timeout 200 python tools/prof_decode.py --layers 4 --shard-shapes 8 > gpurun_out/c9/prof_tp8shapes_mega.txt 2>&1
timeout 200 python tools/prof_decode.py --layers 4 > gpurun_out/c9/prof_tp1_mega.txt 2>&1
tail -n 24 gpurun_out/c9/prof_tp8shapes_mega.txt
