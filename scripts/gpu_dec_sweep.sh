for st in 2 3 4; do for F in 1024 2960; do
echo "stages $st F $F"; FPV_DECODE_STAGES=$st bash scripts/gpu_frames.sh $F; done; done
