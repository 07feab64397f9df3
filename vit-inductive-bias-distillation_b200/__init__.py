"""B200-native BASD distillation-loss hot path.

Drop-in for the reference's ``src.losses`` module API::

    from basd_b200.losses import BASDLoss            # instead of src.losses.combined
"""
from .losses import (BASDLoss, GrassmannianLayerSelector, geometric_relational_loss,  # noqa: F401
                     marchenko_pastur_rank)

__version__ = "0.1.0"
