"""`from_pretrained` / `download_file` (reference hub.py:8-83).  The pretrained weights live on the Hugging
Face hub (`pingzhili/cg-schnet`); with a local path (or an already cached file) the model is loaded, without
network access a clear error is raised instead of a silent fallback."""
import os

DEFAULT_REPO = "pingzhili/cg-schnet"


def download_file(filename: str, repo_id: str = DEFAULT_REPO, **kwargs) -> str:
    if os.path.exists(filename):
        return filename
    try:
        from huggingface_hub import hf_hub_download
        return hf_hub_download(repo_id=repo_id, filename=filename, **kwargs)
    except Exception as err:  # no network / not cached
        raise RuntimeError(f"cannot obtain '{filename}' from '{repo_id}' (offline?): {err}") from err


def from_pretrained(filename: str = "model.pt", repo_id: str = DEFAULT_REPO, device: str = "cpu", **kwargs):
    from .simulation.cli import load_model_file
    return load_model_file(download_file(filename, repo_id, **kwargs)).to(device)
