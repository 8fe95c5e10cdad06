"""Pure-Python model of how k_ingest counts LPMD pairs (metheor_b200/csrc/k_ingest.cu, per-call phase), used to check the
arithmetic itself on CPU against the oracle's restatement of readutil.rs:166-224:

  * every call carries the packed key K = 4 * query index + methylation bit;
  * for a call and one of its anchors (an earlier call of the same read) e = K_call - K_anchor = 4 d + (-1 | 0 | +1), so
    dmin <= d <= dmax  <=>  4 dmin - 1 <= e <= 4 dmax + 1  (one unsigned compare of e - k4lo against k4span), and e is odd
    exactly when the two calls disagree;
  * the two nearest anchors are tested from registers; a third one is only looked for when the second is still within
    dmax, and the walk from there uses plain distances.
Small inputs only (Python loops)."""

U32 = 0xFFFFFFFF
INT32_MAX, INT32_MIN = 2 ** 31 - 1, -2 ** 31


def window_constants(dmin, dmax):
    dmin_c = max(min(dmin, 70000), -70000)
    dmax_c = max(min(dmax, 70000), -70000)
    ok = dmax_c >= dmin_c
    k4lo = 4 * dmin_c - 1 if ok else INT32_MAX
    k4hi = 4 * dmax_c + 1 if ok else INT32_MIN
    k4span = (k4hi - k4lo) if ok else 0
    return k4lo, k4hi, k4span


def read_pairs(rel, meth, dmin, dmax):
    """-> (n_conc, n_disc) of one read the way the kernel counts them."""
    k4lo, k4hi, k4span = window_constants(dmin, dmax)
    K = [4 * r + m for r, m in zip(rel, meth)]
    conc = disc = 0
    for i in range(1, len(K)):
        for j in (1, 2):  # the two nearest anchors, branch-free in the kernel
            if i - j < 0:
                break
            e = K[i] - K[i - j]
            if ((e - k4lo) & U32) <= k4span:
                if e & 1:
                    disc += 1
                else:
                    conc += 1
        if i >= 3 and K[i] - K[i - 2] <= k4hi:  # a third anchor may be in reach: walk on with plain distances
            z = i - 3
            while z >= 0:
                d = rel[i] - rel[z]
                if d > dmax:
                    break
                if d >= dmin:
                    if meth[i] == meth[z]:
                        conc += 1
                    else:
                        disc += 1
                z -= 1
    return conc, disc


def lpmd(soa, dmin, dmax, min_qual):
    """-> (n_conc, n_disc) over all reads of an oracle-style SoA (B.to_oracle_soa) with mapq >= min_qual."""
    off = soa["cpg_off"]
    conc = disc = 0
    for i in range(len(soa["start"])):
        if int(soa["mapq"][i]) < min_qual:
            continue
        a, b = int(off[i]), int(off[i + 1])
        c, d = read_pairs([int(x) for x in soa["cpg_rel"][a:b]], [int(x) for x in soa["cpg_meth"][a:b]], dmin, dmax)
        conc += c
        disc += d
    return conc, disc
