"""`from_pretrained` / `download_file` with the reference's signatures (hub.py:8-83).  The pretrained weights live on the
Hugging Face hub (`pingzhili/cg-schnet`); a `filename` that exists locally is used as is, and without network access (or
without `huggingface_hub`) a clear error is raised instead of a silent fallback."""
import os
from pathlib import Path
from typing import Optional

DEFAULT_REPO = "pingzhili/cg-schnet"


def download_file(repo_id: str = DEFAULT_REPO, filename: str = "1enh_configurations.pt", cache_dir: Optional[str] = None,
                  revision: Optional[str] = None) -> Path:
    if os.path.exists(filename):
        return Path(filename)
    try:
        from huggingface_hub import hf_hub_download
        return Path(hf_hub_download(repo_id=repo_id, filename=filename, cache_dir=cache_dir, revision=revision))
    except Exception as err:  # no network / not cached / package absent
        raise RuntimeError(f"cannot obtain '{filename}' from '{repo_id}' (offline?): {err}") from err


def from_pretrained(repo_id: str = DEFAULT_REPO, filename: str = "model_and_prior.pt", cache_dir: Optional[str] = None,
                    revision: Optional[str] = None):
    from .models import load_and_adapt_old_checkpoint
    return load_and_adapt_old_checkpoint(str(download_file(repo_id, filename, cache_dir, revision)))
