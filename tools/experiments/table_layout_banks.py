"""Shared-memory wavefront counts of assemble_kernel's impulse-response-table accesses (tile stores of phase 5,
row loads of phase 6) for candidate layouts: how He and ldE of SmemLayout (step_kernel.cuh) were chosen."""
# count shared-memory wavefronts of the E-table access patterns for candidate layouts
import itertools
def wavefronts(addrs, width):
    """addrs: list of 32 double-indices (or None for inactive), width in bytes (8 or 16).  8-byte: per half warp, 16-byte: per quarter."""
    group = 16 if width == 8 else 8
    tot = 0
    for g0 in range(0, 32, group):
        lanes = [a for a in addrs[g0:g0 + group] if a is not None]
        if not lanes: continue
        # bank = (byte address / 4) % 32; an access of `width` covers width/4 banks; distinct addresses on the same bank conflict
        per_bank = {}
        for a in lanes:
            for w in range(width // 4):
                bank = (a * 2 + w) % 32
                per_bank.setdefault(bank, set()).add((a * 2 + w) // 32)
        tot += max(len(v) for v in per_bank.values())
    return tot
def frag_row(lane):
    r = lane >> 2
    return (r & 4) | ((r & 1) << 1) | ((r >> 1) & 1)
NY, kNC, p = 3, 6, 100
b_max = 13; b_full = 8
def giant_cols(): return kNC * b_full + 3 * (b_max - b_full)
def analyse(ldE, rho, name):
    # store_e over all tiles
    n_nt = (giant_cols() + 7) // 8
    st = 0; ideal = 0
    for mt in range(NY):
        for nt in range(n_nt):
            for el in range(2):
                addrs = []
                for lane in range(32):
                    a = frag_row(lane); n = 8 * nt + 2 * (lane & 3)
                    if n < kNC * b_full:
                        b, cc = divmod(n, kNC)
                        addrs.append((mt * kNC + cc + el) * ldE + rho(8 * b + a))
                    else:
                        m = n + el - kNC * b_full
                        addrs.append((mt * kNC + 2 * (m % 3)) * ldE + rho(8 * (b_full + m // 3) + a) if n + el < giant_cols() else None)
                st += wavefronts(addrs, 8); ideal += 2
    # gram loads: thread t rows 2t+j, channel ch, delayed shift
    gl = 0; gi = 0
    for warp in range(2):
        for j in range(2):
            for sh in (0, 40):
                addrs = []
                for lane in range(32):
                    r = 2 * (warp * 32 + lane) + j
                    addrs.append(rho(r - sh) if (r < p and r - sh >= 0) else None)
                gl += wavefronts(addrs, 8); gi += 1
    # conv X40: rows r - 39
    cv = 0
    for warp in range(2):
        for j in range(2):
            addrs = []
            for lane in range(32):
                r = 2 * (warp * 32 + lane) + j
                addrs.append(rho(r - 39) if (39 <= r < p) else None)
            cv += wavefronts(addrs, 8)
    print(f"{name:28s} ldE {ldE:4d}: store_e {st} (ideal {ideal}), gram 8B loads per (warp,j,shift) total {gl} over {gi}, conv {cv}")
analyse(106, lambda r: r, "current")
for H in (52, 53, 54, 56, 58, 60, 64):
    for ldE in (2 * H, 2 * H + 2, 2 * H + 4, 2 * H + 6):
        analyse(ldE, lambda r, H=H: (r >> 1) + H * (r & 1), f"split even/odd H={H}")
print("---- general search")
def run(p_, RPT, NY_=3):
    global NY, p, b_max, b_full
    NY, p = NY_, p_
    b_max = (p + 7) // 8
    f = (p - 39 + 7) // 8
    b_full = min(max(f, 1), b_max)
    rows = 8 * b_max
    half = (rows + RPT - 1) // RPT
    best = []
    for hpad in range(0, 9):
        H = half + hpad
        for lpad in range(0, 17, 2):
            ldE = RPT * H + lpad
            rho = lambda r, H=H: (r // RPT) + H * (r % RPT)
            # reuse analyse pieces
            n_nt = (giant_cols() + 7) // 8
            st = 0
            for mt in range(NY):
                for nt in range(n_nt):
                    for el in range(2):
                        addrs = []
                        for lane in range(32):
                            a = frag_row(lane); n = 8 * nt + 2 * (lane & 3)
                            if n < kNC * b_full:
                                b, cc = divmod(n, kNC)
                                addrs.append((mt * kNC + cc + el) * ldE + rho(8 * b + a))
                            else:
                                m = n + el - kNC * b_full
                                addrs.append((mt * kNC + 2 * (m % 3)) * ldE + rho(8 * (b_full + m // 3) + a) if n + el < giant_cols() else None)
                        st += wavefronts(addrs, 8)
            gl = 0
            for warp in range(2):
                for j in range(RPT):
                    for sh in (0, 40, 39):
                        addrs = []
                        for lane in range(32):
                            r = RPT * (warp * 32 + lane) + j
                            addrs.append(rho(r - sh) if (r < p and r - sh >= 0) else None)
                        gl += wavefronts(addrs, 8)
            best.append((st + 5 * gl, st, gl, H, ldE))
    best.sort()
    print(f"p={p_} RPT={RPT} NY={NY_}: rows {rows} half {half}; best (score, store_e, loads, H, ldE):", best[:4], " | unpadded:", [b for b in best if b[3] == half and b[4] == RPT * half][:1])
run(100, 2); run(200, 4); run(100, 2, 2); run(100, 2, 4); run(64, 2); run(41, 2); run(256, 4); run(7, 2)
