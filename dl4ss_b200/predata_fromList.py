"""`from predata_fromList import prepare_data, prepare_datasize` of the reference's scripts
(TDAA_beta/main_run_sstune_EvalVer.py:11): the same generator protocol on the GPU feature path, see compat.py."""
from .compat import prepare_data, prepare_datasize, set_source, get_source, ListFileSource, SyntheticSource  # noqa: F401
