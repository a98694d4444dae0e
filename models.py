"""`from models import GAT` as train.py:18 / train_ppi.py:17 do -- served by the B200 engine."""
from pygat_b200.models import GAT  # noqa: F401
