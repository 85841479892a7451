"""`PYTHONPATH=src` compatibility shim: the implementation lives in cellcomm_b200.intercepts."""
from cellcomm_b200.intercepts import *  # noqa: F401,F403
from cellcomm_b200.intercepts import (DbRecorder, PlotIntercepts, SinkIntercepts,  # noqa: F401
                                      combined_interceptors, offset_iterations, print_losses,
                                      skip_iterations)
