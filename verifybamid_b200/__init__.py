"""verifybamid_b200 -- B200-native contamination-likelihood engine for the VerifyBamID2 hot path.

Only what that one path needs lives here:
  csrc/        hand-written sm_100a kernel + C ABI (libvb2llk.so) + the C++ host side / CLI
  engine.py    ctypes binding of the C ABI (what tests and bench.py call)
  host.py      ctypes binding of the C++ host side (readers, parser, Nelder-Mead) + the CLI path
  problem.py   the flat image of the reference's estimator data the ABI takes
  panels.py    bundled SVD panels, text <-> packed
  synth.py     synthetic contaminated pileups (BASELINE.json configs)
"""
from .problem import PileupProblem  # noqa: F401
from .engine import (LLKEngine, DevicePanel, PeerReduce, VB2Error, VB2_ERR_UNSUPPORTED, eval_many, eval_many_device, context_array, load_library, build_library, device_count,  # noqa: F401
                     pack_host, time_device, time_device_many, time_host, VB2_PANEL_FP32, VB2_PANEL_FP64, VB2_MIN_MAX_DIM, ABI_SYMBOLS)
