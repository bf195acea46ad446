def find_closest_vector_idx(*args, **kwargs):
    raise RuntimeError("centrex_TlF stub: no Hamiltonian available")
