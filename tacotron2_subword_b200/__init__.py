"""tacotron2_subword_b200 -- B200-native (sm_100a) Tacotron2 dual-stream mel decoder behind
the reference's Python API (PhucNguyenAH/tacotron2_subword).  See DESIGN.md."""
from .hparams import create_hparams  # noqa: F401
from .loss_function import Tacotron2Loss  # noqa: F401
from .model import BERT_Tacotron2, Decoder, DropoutReplay, Tacotron2  # noqa: F401

__all__ = ["create_hparams", "BERT_Tacotron2", "Tacotron2", "Decoder", "DropoutReplay", "Tacotron2Loss"]
