def generate_uncoupled_hamiltonian_X(*args, **kwargs):
    raise RuntimeError("centrex_TlF stub: no Hamiltonian available")


def generate_uncoupled_hamiltonian_X_function(*args, **kwargs):
    raise RuntimeError("centrex_TlF stub: no Hamiltonian available")
