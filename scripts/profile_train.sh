# usage (on the GPU box via gpurun): TAG=r1i bash scripts/profile_train.sh
# training-step evidence: bench line, ncu launch list of one step, full capture of the persistent BPTT kernel
set -x
TAG=${TAG:-r1}
mkdir -p gpurun_out
python bench.py --mode train --steps 10 --warmup 3 > gpurun_out/bench_train_${TAG}.json 2> gpurun_out/bench_train_${TAG}.err
python bench.py --mode train --steps 1 --warmup 2 > gpurun_out/plain_train.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_train_${TAG}.csv \
    python bench.py --mode train --steps 1 --warmup 2 > gpurun_out/ncu_train1.log 2>&1
python scripts/trace_rnn_bwd.py 64 lstm tc > gpurun_out/trace_rnn_bwd_${TAG}.txt 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'rnn_bwd_tc' -s 2 -c 1 -f -o gpurun_out/prof_rnn_bwd_${TAG} \
    python scripts/trace_rnn_bwd.py 64 lstm tc > gpurun_out/ncu_train2.log 2>&1
tail -n 2 gpurun_out/ncu_train1.log gpurun_out/ncu_train2.log | cut -c1-200
head -c 300 gpurun_out/bench_train_${TAG}.json
