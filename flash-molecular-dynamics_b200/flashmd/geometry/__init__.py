"""Internal coordinates (reference geometry/internal_coordinates.py:73-223), plain PyTorch: used by the
prior modules' autograd path (CPU / `--disable_optim`); the step path uses fmd_priors_csr."""
import torch


def compute_distance_vectors(pos: torch.Tensor, mapping: torch.Tensor):
    dr = pos[mapping[1]] - pos[mapping[0]]
    d = dr.norm(p=2, dim=1)
    return dr / d[:, None], d


def compute_distances(pos: torch.Tensor, mapping: torch.Tensor, cell_shifts=None) -> torch.Tensor:
    assert mapping.dim() == 2 and mapping.shape[0] == 2
    dr = pos[mapping[1]] - pos[mapping[0]]
    if cell_shifts is not None:
        dr = dr + cell_shifts
    return dr.norm(p=2, dim=1)


def compute_angles_cos(pos: torch.Tensor, mapping: torch.Tensor, cell_shifts=None) -> torch.Tensor:
    assert mapping.dim() == 2 and mapping.shape[0] == 3
    dr1 = pos[mapping[0]] - pos[mapping[1]]
    dr2 = pos[mapping[2]] - pos[mapping[1]]
    return (dr1 * dr2).sum(dim=1) / (dr1.norm(p=2, dim=1) * dr2.norm(p=2, dim=1))


def compute_angles_raw(pos: torch.Tensor, mapping: torch.Tensor, cell_shifts=None) -> torch.Tensor:
    dr1 = pos[mapping[0]] - pos[mapping[1]]
    dr2 = pos[mapping[2]] - pos[mapping[1]]
    cross = torch.cross(dr1, dr2, dim=1).norm(p=2, dim=1)
    return torch.atan2(cross, (dr1 * dr2).sum(dim=1))


def compute_angles(pos, mapping, cell_shifts=None):
    return compute_angles_cos(pos, mapping, cell_shifts)


def compute_torsions(pos: torch.Tensor, mapping: torch.Tensor, cell_shifts=None) -> torch.Tensor:
    """Dihedral angle with the MDTraj sign convention (internal_coordinates.py:174-223)."""
    assert mapping.dim() == 2 and mapping.shape[0] == 4
    unit = torch.nn.functional.normalize
    b1 = unit(pos[mapping[1]] - pos[mapping[0]], dim=1)
    b2 = unit(pos[mapping[2]] - pos[mapping[1]], dim=1)
    b3 = unit(pos[mapping[3]] - pos[mapping[2]], dim=1)
    n1 = torch.cross(b1, b2, dim=1)
    n2 = torch.cross(b2, b3, dim=1)
    m1 = torch.cross(n1, b2, dim=1)
    return torch.atan2(-(m1 * n2).sum(-1), (n1 * n2).sum(-1))
