"""`PYTHONPATH=src` compatibility shim (reference runtests:3): the implementation lives in
cellcomm_b200.cell_type_training."""
from cellcomm_b200.cell_type_training import *  # noqa: F401,F403
from cellcomm_b200.cell_type_training import __dict__ as _d

globals().update({k: v for k, v in _d.items() if not k.startswith('__')})
