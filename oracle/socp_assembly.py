"""TEST INFRASTRUCTURE — numpy restatement of how the reference assembles the trajectory-dependent SOCP rows.

Follows, triplet by triplet, `build_model` of the reference (`rocketland.jl:117-133` dynamics rows, `194-201` thrust lower
bound) and its refresh in `solve_step` (`251-265`), which writes the same coefficients.  MathOptInterface sums duplicate
(row, variable) terms of a VectorAffineFunction, so the dense matrix is accumulated with `+=`.

PARITY UNPINNED: the reference has no fixture for these rows and cannot run here (no Julia / MOI / Mosek); the
restatement is anchored on the call sites cited above.  Only tests/ may import this module.
"""
import numpy as np


def variable_index(K):
    """0-based indices in the reference's creation order xv, uv, dxv, duv, dsig, nuv (rocketland.jl:71-76), as (dim, K+1)
    arrays filled column-major like Julia's `reshape(add_variables(...), dim, K+1)`."""
    n = K + 1
    pos = 0
    out = {}
    for name, dim in (("xv", 14), ("uv", 3), ("dxv", 14), ("duv", 3)):
        out[name] = pos + np.arange(dim * n).reshape(n, dim).T
        pos += dim * n
    out["dsig"] = pos
    pos += 1
    out["nuv"] = pos + np.arange(14 * n).reshape(n, 14).T
    pos += 14 * n
    out["n_vars"] = pos
    return out


def assemble_dense(derivatives, endpoints, states, controls, Tmin):
    """derivatives (K, 14, 21), endpoints (K, 14), states (K+1, 14), controls (K+1, 3)  ->
    (M, const): M is (15K+1) x n_vars over the reference's own variable numbering, const the VAF constants."""
    K = derivatives.shape[0]
    v = variable_index(K)
    M = np.zeros((15 * K + 1, v["n_vars"]))
    const = np.zeros(15 * K + 1)
    for n in range(K):                                            # rocketland.jl:124  for n=1:K
        variables = np.concatenate([v["dxv"][:, n], v["duv"][:, n], v["duv"][:, n + 1], [v["dsig"]]])   # :126
        for col, var in zip(derivatives[n].T, variables):         # :125  eachcol(derivative) paired with the variables
            for i, x in enumerate(col):                           #       enumerate(map(x -> SA(x, var), col))
                M[14 * n + i, var] += x
        for i in range(14):
            M[14 * n + i, v["nuv"][i, n + 1]] += 1.0              # :127
            M[14 * n + i, v["dxv"][i, n + 1]] += -1.0             # :128
        const[14 * n:14 * n + 14] = endpoints[n] - states[n + 1]  # :129
    for n in range(K + 1):                                        # :199-200
        u = controls[n, :3]
        nu = np.sqrt(u[0] * u[0] + u[1] * u[1] + u[2] * u[2])
        for j in range(3):
            M[14 * K + n, v["duv"][j, n]] += -(u[j] / nu)
        const[14 * K + n] = Tmin - nu
    return M, const
