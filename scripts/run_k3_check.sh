# GPU box: parity + timing + per-phase trace of the tcgen05 recurrent kernel (every step under its own timeout)
timeout 60 python scripts/debug_rnn_hang.py 3 5 2>&1 | head -1
timeout 120 python scripts/check_rnn_tc.py 2>&1 | tail -9
timeout 120 python scripts/time_encode.py 2>&1 | tail -9
timeout 60 python scripts/trace_rnn_tc.py 256 2>&1 | tail -20
DL4SS_RNN_TILES_PER_CTA=3 timeout 60 python scripts/trace_rnn_tc.py 256 2>&1 | tail -20
