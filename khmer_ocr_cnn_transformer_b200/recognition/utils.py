"""Logging setup and checkpoint shape sniffing
(reference: netra_ocr/recognition/utils.py:7-43)."""
import logging
from pathlib import Path

logger = logging.getLogger(__name__)


def setup_logging():
    logging.basicConfig(level=logging.INFO, format="%(asctime)s - %(levelname)s - %(message)s",
                        datefmt="%H:%M:%S")


def autodetect_config(model_path) -> dict:
    """Infer {max_seq_len, emb_dim, decode_max_len} from the checkpoint's `global_pos` and
    `dec.pos_emb` shapes.  Raises FileNotFoundError for a missing file, like the reference."""
    from . import _core
    load_checkpoint = _core.checkpoint.load_checkpoint
    path = Path(model_path)
    if not path.exists():
        raise FileNotFoundError(f"Model not found at {path}")
    logger.info(f"Inspecting checkpoint: {path.name}...")
    sd = load_checkpoint(path)
    detected = {}
    if "global_pos" in sd:
        shape = sd["global_pos"].shape
        detected["max_seq_len"] = int(shape[0])
        detected["emb_dim"] = int(shape[1])
    if "dec.pos_emb" in sd:
        detected["decode_max_len"] = int(sd["dec.pos_emb"].shape[0])
    return detected
