from .torch_impl import radius_graph, radius_graph_csr, torch_neighbor_list  # noqa: F401
from .neighbor_list import atomic_data2neighbor_list, make_neighbor_list, validate_neighborlist  # noqa: F401
