"""Drop-in mirror of the reference's ``src/losses`` package (same names, same signatures)."""
from .combined import BASDLoss, _align_token_count
from .layer_selector import GrassmannianLayerSelector, marchenko_pastur_rank
from .relational import geometric_relational_loss

__all__ = ["BASDLoss", "GrassmannianLayerSelector", "geometric_relational_loss",
           "marchenko_pastur_rank", "_align_token_count"]
