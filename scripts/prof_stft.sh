set -x
TAG=${TAG:-r1c}
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-train-extra"
mkdir -p gpurun_out
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'stft256|istft' -s 6 -c 2 -o gpurun_out/prof_stft_${TAG} $CMD > gpurun_out/ncu2.log 2>&1
tail -n 2 gpurun_out/ncu2.log
