"""Site types -- restates /root/reference/src/sitetypes.jl:1-166 and
/root/reference/src/lattices/spinhalf.jl:1-23 (Pauli convention, eigenvalues +-1)."""
import numpy as np


class Sitetypes:
    """sitetypes.jl:1-9.  Named local states and operators with their daggers."""

    def __init__(self, dim):
        self.dim = dim
        self.statenames, self.states = [], []
        self.opnames, self.ops, self.opdags = [], [], []
        self.temp = 0

    def add_state(self, name, vec):  # sitetypes.jl:85-96
        vec = np.asarray(vec, dtype=np.complex128)
        if vec.shape != (self.dim,):
            raise ValueError(f"The vector must be dimension {self.dim}.")
        if name in self.statenames:
            raise ValueError(f"The state name {name} already exists.")
        self.statenames.append(name)
        self.states.append(vec)

    def add_op(self, name, mat, dag="None"):  # sitetypes.jl:105-120
        mat = np.asarray(mat, dtype=np.complex128)
        if mat.shape != (self.dim, self.dim):
            raise ValueError(f"The matrix must be dimensions ({self.dim}, {self.dim}).")
        if name in self.opnames:
            raise ValueError(f"The operator name {name} already exists.")
        self.opnames.append(name)
        self.ops.append(mat)
        self.opdags.append(dag)

    def state(self, name):  # sitetypes.jl:27-37
        if name not in self.statenames:
            raise KeyError(f"The state {name} is undefined.")
        return self.states[self.statenames.index(name)].copy()

    def op(self, name):  # sitetypes.jl:45-55
        if name not in self.opnames:
            raise KeyError(f"The operator {name} is undefined.")
        return self.ops[self.opnames.index(name)].copy()

    def dag(self, name):  # sitetypes.jl:62-72
        if name not in self.opnames:
            raise KeyError(f"The operator {name} is undefined.")
        return self.opdags[self.opnames.index(name)]

    def opprod(self, names):  # sitetypes.jl:144-166
        prod = None
        for nm in names:
            prod = self.op(nm) if prod is None else prod @ self.op(nm)
        for i, o in enumerate(self.ops):
            if np.sum(np.abs(prod - o)) < 1e-12:
                return self.opnames[i]
        name, dagname = f"temp{self.temp}", f"temp{self.temp + 1}"
        self.add_op(name, prod, dagname)
        self.add_op(dagname, prod.conj().T, name)
        self.temp += 2
        return name


def spinhalf():
    """lattices/spinhalf.jl:1-23."""
    st = Sitetypes(2)
    st.add_state("up", [1, 0])
    st.add_state("dn", [0, 1])
    st.add_state("s", [0.5 ** 0.5, 0.5 ** 0.5])
    st.add_state("as", [0.5 ** 0.5, -0.5 ** 0.5])
    st.add_op("x", [[0, 1], [1, 0]], "x")
    st.add_op("y", [[0, -1j], [1j, 0]], "y")
    st.add_op("z", [[1, 0], [0, -1]], "z")
    st.add_op("id", [[1, 0], [0, 1]], "id")
    st.add_op("pu", [[1, 0], [0, 0]], "pu")
    st.add_op("pd", [[0, 0], [0, 1]], "pd")
    st.add_op("n", [[1, 0], [0, 0]], "n")
    st.add_op("s+", [[0, 1], [0, 0]], "s-")
    st.add_op("s-", [[0, 0], [1, 0]], "s+")
    return st
