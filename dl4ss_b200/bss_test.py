"""`import bss_test; bss_test.cal(path, aim_mix_number)` of the reference's scripts (Torch_multi/bss_test.py:12-61):
SDR of the wav files `bss_eval` wrote (or of its in-memory batch), scored on the device, see compat.py."""
from .compat import bss_test as _impl

add_slience_channel = _impl.add_slience_channel
cal = _impl.cal
