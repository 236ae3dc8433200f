"""B200-native SOM layer hot path for ViT-SOM (sm_100a, tcgen05/TMA), behind the reference's SOMLayer API.

    from vit_som_b200 import SOMLayer      # drop-in for /root/reference/models/som_layer.py:SOMLayer

The numeric work lives in ``libsom_b200.so`` (C ABI: include/som_b200.h); build it with
``python -c "import __graft_entry__ as g; g.build()"``.  There is no CPU path and no fallback.
"""
from ._lib import SomError, build, lib  # noqa: F401
from .glue import gamma_ramp, som_input  # noqa: F401
from .graph import StepGraph  # noqa: F401
from .optim import FusedPrototypeAdamW  # noqa: F401
from .som_layer import NeighbourhoodWeights, SOMLayer  # noqa: F401

__all__ = ["SOMLayer", "NeighbourhoodWeights", "FusedPrototypeAdamW", "StepGraph", "gamma_ramp", "som_input", "SomError",
           "build", "lib"]
