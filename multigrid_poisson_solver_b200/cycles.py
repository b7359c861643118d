"""Generators for Cycle.txt cycle-structure files.

Format (reference README.md:43-128, parser MG_solver_CPU.cpp:103-109,160,171-189,307,331-344):
  line 1  L min_x min_y
  line 2  con_step con_N      con_step: -1 error trigger, 0 per-node, k fixed sweeps
                              con_N:     0 per-node next_N, 1 halve, 2 minus one
  line 3  N_max N_min
  then the node stream: -1 (smooth+restrict), 0 tol option (exact solve),
  1 (prolong+smooth), 2 (stop), each followed by the options its mode needs.
The shipped files end with "2" and no trailing newline; so do these.
"""


def _fmt_tol(tol):
    s = ("%.12f" % tol).rstrip("0")
    return s if not s.endswith(".") else s + "0"


def ladder(N_max, N_min, con_N=1):
    """Grid sizes the reference builds (MG_solver_CPU.cpp:111-146)."""
    if con_N == 1:
        out, n = [], N_max
        while n >= N_min:
            out.append(n)
            n //= 2
        return out
    if con_N == 2:
        return list(range(N_max, N_min - 1, -1))
    raise ValueError("con_N must be 1 or 2 for an automatic ladder")


def _header(L, min_x, min_y, con_step, con_N, N_max, N_min):
    return ["%s %s %s" % (repr(float(L)), repr(float(min_x)), repr(float(min_y))),
            "%d %d" % (con_step, con_N), "%d %d" % (N_max, N_min)]


def v_cycle(N_max, N_min, step=3, tol=1e-7, option=1, con_N=1, L=1.0, min_x=0.0, min_y=0.0, cycles=1):
    """V-cycle down the whole ladder (src/Vcycle.txt is v_cycle(256, 8));
    step=-1 gives src/VcycleTrigger.txt's shape.  cycles>1 chains cycles through
    the init==0 restart path (MG_solver_CPU.cpp:209-211)."""
    depth = len(ladder(N_max, N_min, con_N)) - 1
    lines = _header(L, min_x, min_y, step, con_N, N_max, N_min)
    for _ in range(cycles):
        lines += ["-1"] * depth + ["0", "%s %d" % (_fmt_tol(tol), option)] + ["1"] * depth
    lines.append("2")
    return "\n".join(lines)


def w_cycle(N_max, N_min, levels=None, step=3, tol=1e-8, option=1, con_N=1, L=1.0, min_x=0.0, min_y=0.0):
    """W-cycle with the recursion of src/Wcycle.txt (= w_cycle(256, 8, levels=3)):
    top = -1, W(1), 1;  W(l) = coarsest ? '0 tol opt' : -1, W(l+1), 1, -1, W(l+1), 1.
    `levels` = number of restrictions from the top (default: the whole ladder)."""
    depth = len(ladder(N_max, N_min, con_N)) - 1
    if levels is None:
        levels = depth
    assert 1 <= levels <= depth
    exact = ["0", "%s %d" % (_fmt_tol(tol), option)]

    def w(l):
        if l == levels:
            return list(exact)
        inner = w(l + 1)
        return ["-1"] + inner + ["1", "-1"] + inner + ["1"]

    lines = _header(L, min_x, min_y, step, con_N, N_max, N_min)
    lines += ["-1"] + w(1) + ["1", "2"]
    return "\n".join(lines)


def two_grid(N_max, N_min, step=3, tol=1e-8, option=1):
    """src/test.txt's shape: one restriction, exact solve, one prolongation."""
    lines = _header(1.0, 0.0, 0.0, step, 1, N_max, N_min)
    lines += ["-1", "0", "%s %d" % (_fmt_tol(tol), option), "1", "2"]
    return "\n".join(lines)


def manual(nodes, N_max, N_min, L=1.0, min_x=0.0, min_y=0.0, con_step=0, con_N=0):
    """Free-form stream: `nodes` is a list of tuples
    (-1, step, next_N) / (0, tol, option) / (1, step); fields that the chosen
    con_step/con_N mode does not read are dropped."""
    lines = _header(L, min_x, min_y, con_step, con_N, N_max, N_min)
    for nd in nodes:
        if nd[0] == -1:
            lines.append("-1")
            opts = []
            if con_step == 0:
                opts.append(str(nd[1]))
            if con_N == 0:
                opts.append(str(nd[2]))
            if opts:
                lines.append(" ".join(opts))
        elif nd[0] == 0:
            lines += ["0", "%s %d" % (_fmt_tol(nd[1]), nd[2])]
        elif nd[0] == 1:
            lines.append("1")
            if con_step == 0:
                lines.append(str(nd[1]))
    lines.append("2")
    return "\n".join(lines)


def tokens(text):
    """Whitespace-delimited numeric tokens, the way `ifstream >>` sees the file."""
    return [float(t) for t in text.split()]
