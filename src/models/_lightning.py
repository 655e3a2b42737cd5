"""`pl.LightningModule` when pytorch_lightning is installed (real deployments), otherwise a minimal stand-in with
the attributes the hot path touches (`device`, `global_rank`, `logger`), so the package imports on a bare box."""
import torch
import torch.nn as nn

try:  # pragma: no cover - not installed in the build container
    import pytorch_lightning as pl
    LightningModule = pl.LightningModule
    HAVE_LIGHTNING = True
except Exception:
    HAVE_LIGHTNING = False

    class LightningModule(nn.Module):
        global_rank = 0
        logger = None

        @property
        def device(self):
            p = next(self.parameters(), None)
            if p is not None:
                return p.device
            b = next(self.buffers(), None)
            return b.device if b is not None else torch.device("cpu")
