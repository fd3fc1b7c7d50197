"""rectipy_b200 -- B200-native (sm_100a) engine for RectiPy's time-stepped integration hot path.

Drop-in for the reference's public surface on that path: `Network` / `FeedbackNetwork` (add_diffeq_node, add_func_node, add_edge,
compile, run, fit_bptt, fit_ridge, fit_rls, test), `Observer`, the node/edge classes and the connectivity helpers.
All numerical work is done by hand-written CUDA kernels behind a C ABI (include/rectipy_b200.h); there is no CPU
fallback.
"""
from .network import Network
from .feedback import FeedbackNetwork
from .observer import Observer
from .nodes import RateNet, SpikeResetNet, InstantNode
from .edges import Linear, LinearFilter, LinearMasked, LinearMemory, LinearMemoryFilter, RLS
from .utility import (random_connectivity, circular_connectivity, line_connectivity, input_connections, normalize,
                      wta_score, readout)
from . import engine, templates, parallel

__version__ = "0.1.0"
__all__ = ["Network", "FeedbackNetwork", "Observer", "RateNet", "SpikeResetNet", "InstantNode", "Linear", "LinearMasked", "LinearMemory", "LinearFilter", "LinearMemoryFilter", "RLS",
           "random_connectivity", "circular_connectivity", "line_connectivity", "input_connections", "normalize",
           "wta_score", "readout", "engine", "templates", "parallel"]
