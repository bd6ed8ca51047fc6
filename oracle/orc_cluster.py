"""oracle/ -- TEST INFRASTRUCTURE ONLY (see oracle/orc_common.h for the rules of this directory).

Literal pure-Python restatement of the aggregation and clustering modules of `spacedust clustersearch`
(small inputs only): prefixid, besthitbyset, mergeresultsbyset, combinehits, clusterhits, summarizeresults.
Every function works on the text records the reference modules exchange and cites the reference lines it
follows.  Parity pin: tests/test_oracle_cpu.py diffs each stage against the DBs written by the unmodified
reference binary (oracle/_ref/spacedust)."""
import math
import sys

DBL_MIN = sys.float_info.min
DBL_MAX = sys.float_info.max


def fmt3e(x):
    """printf("%.3E") / fmt "{:.3E}" (Util.cpp:658-660)."""
    if x != x:
        return "NAN"
    if x in (float("inf"), float("-inf")):
        return "INF" if x > 0 else "-INF"
    return "%.3E" % x


# ------------------------------------------------------------------------------------------------ prefixid
def prefixid(result):
    """lib/mmseqs/src/util/prefixid.cpp:12-84: prepend the entry key to every line.  result: {key: [line, ...]}"""
    return {k: ["%d\t%s" % (k, l) for l in lines] for k, lines in result.items()}


# ------------------------------------------------------------------------------------------------ besthitbyset
def compute_log_pval(evalue, log_calibration=0.0):
    """src/util/besthitbyset.cpp:10-20"""
    if evalue == 0:
        return math.log(DBL_MIN) - log_calibration
    if 0 < evalue < 10e-4:
        return math.log(evalue) - log_calibration
    return math.log(1 - math.exp(-evalue)) - log_calibration


def besthitbyset(result_prefixed, member_to_set, simple_best_hit=True, subopt_factor=0):
    """src/util/besthitbyset.cpp:41-144 + Aggregation::run (src/util/Aggregation.cpp:24-91).
    result_prefixed: {geneKey: [line]}, line columns: qKey tKey bits seqId eval ...; -> {geneKey: [line]}"""
    assert simple_best_hit and subopt_factor == 0, "only the clustersearch defaults are restated"
    out = {}
    for key, lines in result_prefixed.items():
        groups = {}
        for line in lines:
            if not line:
                continue
            cols = line.split("\t")
            groups.setdefault(member_to_set[int(cols[1])], []).append(cols)
        res = []
        for set_key in sorted(groups):
            best_eval, best = DBL_MAX, None
            for cols in groups[set_key]:
                ev = float(cols[4])
                if ev < best_eval:
                    best_eval, best = ev, cols
            if best is None:
                res.append("")
                continue
            cols = list(best)
            cols[2] = fmt3e(compute_log_pval(best_eval))
            res.append("\t".join(cols))
        out[key] = res
    return out


# ------------------------------------------------------------------------------------------------ mergeresultsbyset
def mergeresultsbyset(set_to_member, aggregate):
    """lib/mmseqs/src/util/mergeresultsbyset.cpp:11-71: concatenate the entries of a set's members."""
    out = {}
    for set_key, members in set_to_member.items():
        lines = []
        for m in members:
            lines.extend(aggregate.get(m, []))
        out[set_key] = lines
    return out


# ------------------------------------------------------------------------------------------------ combinehits
def precompute_log_b(orf_count, pval_thr, lgamma_lookup):
    """lib/mmseqs/src/multihit/combinepvalperset.cpp:12-26"""
    log_thr = math.log(pval_thr)
    log1m = math.log(1 - pval_thr)
    logb = [0.0] * orf_count
    logb[orf_count - 1] = orf_count * log_thr
    for i in range(orf_count - 2, -1, -1):
        k = i + 1
        lbin = lgamma_lookup[orf_count + 1] - lgamma_lookup[orf_count - k + 1] - lgamma_lookup[k + 1]
        new_term = lbin + k * log_thr + (orf_count - k) * log1m
        logb[i] = logb[i + 1] + math.log(1 + math.exp(new_term - logb[i + 1]))
    return logb


def combinehits(aggregate_merged, member_to_set_t, q_set_size, t_set_size, alpha=1.0, filter_self_match=False):
    """src/util/combinehits.cpp:74-234 (aggregation-mode 0) + Aggregation::runWithHeader (Aggregation.cpp:93-157).
    -> list of (header line, [hit lines]) in (query set asc, target set asc) order; keys are positional."""
    max_orf = max(q_set_size.values())
    lgl = [math.lgamma(i) if i > 0 else float("inf") for i in range(max_orf + 2)]
    n_target_sets = len(t_set_size)
    out = []
    for q_set in sorted(aggregate_merged):
        groups = {}
        for line in aggregate_merged[q_set]:
            if not line:
                continue
            cols = line.split("\t")
            groups.setdefault(member_to_set_t[int(cols[1])], []).append(cols)
        orf = q_set_size[q_set]
        logb = precompute_log_b(orf, alpha / (orf + 1), lgl)
        for t_set in sorted(groups):
            if filter_self_match and q_set == t_set:
                continue
            pthr = 10e-7
            log_thr = math.log(pthr)
            k, r, entries = 0, 0.0, []
            for cols in groups[t_set]:
                lp = float(cols[2])
                if lp < log_thr:
                    k += 1
                    r -= lp - log_thr
                    entries.append(cols)
            if r == 0 or k == 0:
                continue
            hdr = "%d\t%d\t%d\t%d\t%d\t" % (q_set, t_set, orf, t_set_size[t_set], k)
            if math.isinf(r):
                hdr += fmt3e(0.0)
            else:
                emr = math.exp(-r)
                if emr == 0:
                    hdr += fmt3e(0.0)
                else:
                    logr = math.log(r)
                    s = 0.0
                    for i in range(orf):
                        try:
                            s += math.exp(i * logr - lgl[i + 1] + logb[i])
                        except OverflowError:
                            s += float("inf")
                    hdr += fmt3e(emr * s * n_target_sets)
            body = []
            for cols in entries:
                c = list(cols)
                c[2] = fmt3e(math.exp(float(c[2])))
                body.append("\t".join(c))
            out.append((hdr, body))
    return out


# ------------------------------------------------------------------------------------------------ clusterhits
GAMMA_R10 = 10.900511
GAMMA_DK = [2.48574089138753565546e-5, 1.05142378581721974210, -3.45687097222016235469, 4.51227709466894823700,
            -2.98285225323576655721, 1.05639711577126713077, -1.95428773191645869583e-1, 1.70970543404441224307e-2,
            -5.71926117404305781283e-4, 4.63399473359905636708e-6, -2.71994908488607703910e-9]
GAMMA_C = 2 * math.sqrt(math.exp(1.0) / math.pi)


def log_gamma(x):
    """src/util/ClusterHits.cpp:42-63 (Lanczos; note the integer abs() on the reflection branch, :44)."""
    if x < 0.5:
        return math.log(math.pi) - math.log(abs(int(math.sin(math.pi * x)))) - log_gamma(1 - x) if int(math.sin(math.pi * x)) != 0 \
            else float("inf")
    if x == 1:
        return 0.0
    s = GAMMA_DK[0]
    for i in range(1, 11):
        s += GAMMA_DK[i] / (x + (i - 1))
    return math.log(GAMMA_C) + (x - 0.5) * math.log(x + GAMMA_R10 - 0.5) - (x - 0.5) + math.log(s)


def _u32(x):
    return x & 0xFFFFFFFF


class Hit(object):
    __slots__ = ("alignment", "pval", "qPos", "tPos", "qStrand", "tStrand")


def log_cluster_pval(lookup, k, m, q0=0.001):
    return 2 * lookup[m + 1] - 2 * lookup[m - k + 1] - lookup[k + 1] + k * math.log(q0)


def log_ordering_pval(lookup, k, m):
    return math.log(1 - 1.0 * m / k) - m * math.log(2) - lookup[m + 1]


def find_span(cluster):
    i_max = max(h.qPos for h in cluster)
    i_min = min(h.qPos for h in cluster)
    j_max = max(h.tPos for h in cluster)
    j_min = min(h.tPos for h in cluster)
    return max(i_max - i_min + 1, j_max - j_min + 1)


def find_conserved_pairs(cluster):
    cluster.sort(key=lambda h: h.qPos)          # qPos is unique inside a genome pair
    m = 0
    for l in range(len(cluster) - 1):
        same_order = cluster[l + 1].tPos > cluster[l].tPos
        s1 = cluster[l].qStrand == cluster[l].tStrand
        s2 = cluster[l + 1].qStrand == cluster[l + 1].tStrand
        if s1 == same_order and s2 == same_order:
            m += 1
    return m


def cluster_match_score(lookup, cluster):
    if not cluster:
        return 0.0
    span = find_span(cluster)
    k = len(cluster)
    m = find_conserved_pairs(cluster)
    return -0.5 * log_cluster_pval(lookup, k, span) - 0.5 * log_ordering_pval(lookup, k, m)


def is_compatible(c1, c2, d):
    """ClusterHits.cpp:137-159: the differences are unsigned, overlapping ranges wrap around and fail."""
    i_max1 = max(h.qPos for h in c1); i_min1 = min(h.qPos for h in c1)
    j_max1 = max(h.tPos for h in c1); j_min1 = min(h.tPos for h in c1)
    i_max2 = max(h.qPos for h in c2); i_min2 = min(h.qPos for h in c2)
    j_max2 = max(h.tPos for h in c2); j_min2 = min(h.tPos for h in c2)
    return min(_u32(j_min1 - j_max2), _u32(j_min2 - j_max1)) <= d and min(_u32(i_min1 - i_max2), _u32(i_min2 - i_max1)) <= d


def group_nodes(nodes, match, i, j, d):
    if nodes[i] and nodes[j]:
        c1 = [match[x] for x in nodes[i]]
        c2 = [match[x] for x in nodes[j]]
        if is_compatible(c1, c2, d):
            return c1 + c2
    return []


def multihit_pval(lookup, cluster, nq, alpha):
    k, r = 0, 0.0
    log_thr = math.log(alpha / (nq + 1))
    for h in cluster:
        lp = math.log(h.pval) if h.pval > 0 else float("-inf")
        if lp < log_thr:
            k += 1
            r -= lp - log_thr
    if r == 0:
        return 1.0
    if math.isinf(r):
        return 0.0
    emr = math.exp(-r)
    if emr == 0:
        return 0.0
    s = 0.0
    for i in range(k - 1):
        s += math.pow(r, i) / math.exp(lookup[i + 1])
    return emr * s


def parse_lookup_name(name):
    """`<acc>_<geneIdx>_<start>_<end>` (data/createsetdb.sh:121-135) -> (pos, strand) as ClusterHits.cpp:328-350 reads it."""
    cols = name.split("_")
    start, end = int(cols[-2]), int(cols[-1])
    return int(cols[-3]), 1 if start < end else 0


def clusterhits(matches, q_lookup, t_lookup, max_orf, d=3, cls=2, p_clu_thr=0.01, p_mh_thr=0.01, alpha=1.0):
    """src/util/ClusterHits.cpp:215-510.  matches: list of (header, [hit lines]); lookups: {key: name}.
    -> list of (header line, [alignment lines]) in output order (genome pair order, node order)."""
    lookup = [log_gamma(float(i)) if i > 0 else float("inf") for i in range(max_orf + 2)]
    out = []
    for hdr, body in matches:
        hc = hdr.split("\t")
        q_set, t_set, nq = int(hc[0]), int(hc[1]), int(hc[2])
        match = []
        for line in body:
            cols = line.split("\t")
            h = Hit()
            h.alignment = line
            h.pval = float(cols[2])
            h.qPos, h.qStrand = parse_lookup_name(q_lookup[int(cols[0])])
            h.tPos, h.tStrand = parse_lookup_name(t_lookup[int(cols[1])])
            match.append(h)
        K = len(match)
        if K == 1:
            continue
        dist = [[0.0] * K for _ in range(K)]
        dmin = [0] * K
        nodes = [[n] for n in range(K)]
        for i in range(K):
            for j in range(K):
                if i != j:
                    dist[i][j] = cluster_match_score(lookup, group_nodes(nodes, match, i, j, d))
                if dist[i][j] > dist[i][dmin[i]]:
                    dmin[i] = j
        s_min = -0.5 * log_cluster_pval(lookup, 2, d + 1) - 0.5 * log_ordering_pval(lookup, 2, 1)
        max_score = DBL_MAX
        first = True
        while first or max_score >= s_min:
            i1 = 0
            for i in range(K):
                if dist[i][dmin[i]] > dist[i1][dmin[i1]]:
                    i1 = i
            i2 = dmin[i1]
            max_score = dist[i1][i2]
            if max_score != 0:
                first = False
            else:
                break
            nodes[i1].extend(nodes[i2])
            nodes[i2] = []
            for j in range(K):
                if i1 == j or i2 == j:
                    dist[i1][j] = 0.0
                    dist[j][i1] = 0.0
                else:
                    dist[i1][j] = cluster_match_score(lookup, group_nodes(nodes, match, i1, j, d))
                    dist[j][i1] = dist[i1][j]
                dist[i2][j] = 0.0
                dist[j][i2] = 0.0
                if j != 0:
                    if dist[i1][j] > dist[i1][dmin[i1]]:
                        dmin[i1] = j
                else:
                    dmin[i1] = j
                if j != i1 and j != i2:
                    if dist[j][i1] > dist[j][dmin[j]]:
                        dmin[j] = i1
        for i in range(K):
            if len(nodes[i]) >= cls:
                cluster = [match[x] for x in nodes[i]]
                p_co = math.exp(-cluster_match_score(lookup, cluster))
                p_mh = multihit_pval(lookup, cluster, nq, alpha)
                if p_co <= p_clu_thr and p_mh <= p_mh_thr:
                    out.append(("%d\t%d\t%s\t%s\t%d" % (q_set, t_set, fmt3e(p_co), fmt3e(p_mh), len(cluster)),
                                [h.alignment for h in cluster]))
    return out
