from .torch_impl import radius_graph, radius_graph_csr  # noqa: F401
