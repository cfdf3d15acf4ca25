# usage (on the GPU box via gpurun): TAG=r2c bash scripts/profile.sh
# bench line, ncu launch list of the same command, full captures of the HBM kernels and of the recurrent / projection kernels
set -x
TAG=${TAG:-r2}
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-train-extra"
mkdir -p gpurun_out
timeout 600 python bench.py > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
timeout 200 $CMD > gpurun_out/plain.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_${TAG}.csv $CMD > gpurun_out/ncu1.log 2>&1
# full captures: eager launches, one batch at a time (--graph 0) so that -s / -c count the same kernels whatever the graph warm-up does
timeout 200 $CMD --graph 0 > gpurun_out/plain2.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'stft256|istft' -s 6 -c 2 -f -o gpurun_out/prof_stft_${TAG} $CMD --graph 0 > gpurun_out/ncu2.log 2>&1
timeout 200 $CMD --graph 0 > gpurun_out/plain3.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'rnn_tc|gemm2?_bf16x3' -s 12 -c 4 -f -o gpurun_out/prof_rnn_${TAG} $CMD --graph 0 > gpurun_out/ncu3.log 2>&1
tail -n 2 gpurun_out/ncu1.log gpurun_out/ncu2.log gpurun_out/ncu3.log
head -c 400 gpurun_out/bench_${TAG}.json
