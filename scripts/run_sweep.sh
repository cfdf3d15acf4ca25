# BASELINE configs[4] on the GPU box: hop {128, 64} x (S=2 real, S=3 cRM); NG GPUs (default 1)
NG=${NG:-1}
OUT=${OUT:-gpurun_out/sweep_stft_r2_${NG}gpu.md}
PTS=${PTS:-}
RUN="python"
if [ "$NG" != "1" ]; then RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29511"; fi
echo "# STFT / mask+iSTFT sweep (BASELINE configs[4]), $NG B200" > $OUT; echo >> $OUT
for cfg in "--hop 128 --S 2" "--hop 64 --S 2" "--hop 128 --S 3 --crm" "--hop 64 --S 3 --crm"; do
  timeout 400 $RUN scripts/sweep_stft.py $cfg $PTS --out $OUT --append > gpurun_out/sweep_log_${NG}.txt 2>&1 || tail -5 gpurun_out/sweep_log_${NG}.txt
done
cat $OUT
