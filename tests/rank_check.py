"""Node-id retrieval contract, in numbers.  The ranking kernel is bit-exact on a given distribution
(test_rank_candidates_matches_reference_lists); end to end, two candidates can only change places where the REFERENCE
itself separates them by less than ``margin`` relative (2e-5: the accuracy of an fp32 forward; the reference is only
~1e-9-stable under its own per-batch fact shuffle).  ``compare`` asserts that and returns the counts, which
every caller prints and appends to gpurun_out/ranking_swaps.jsonl so the numbers of a GPU run can be read afterwards."""
import json
import os


def compare(got, ref, ref_dist, margin=2e-5):
    """got: list of evaluate.Retrieved; ref: oracle/reference lists [(node, entity, prob), ...] per question;
    ref_dist: the reference pred_dist [B, N].  -> dict(positions, swaps, cut_moves, questions_exact, max_swap_gap)."""
    positions = swaps = cut_moves = exact = 0
    max_gap = 0.0
    for b, (r, rr) in enumerate(zip(got, ref)):
        gi = r.idx.tolist()
        ri = [n for n, _, _ in rr]
        positions += len(ri)
        if gi == ri:
            exact += 1
            continue
        assert abs(len(gi) - len(ri)) <= 1, (b, len(gi), len(ri))      # cut may move by one near-tied item
        cut_moves += int(len(gi) != len(ri))
        for i in range(min(len(gi), len(ri))):
            if gi[i] != ri[i]:
                p_here = rr[i][2]
                gap = abs(float(ref_dist[b][gi[i]]) - p_here) / p_here
                assert gap <= margin, (b, i, gi[i], ri[i], gap)
                max_gap = max(max_gap, gap)
                swaps += 1
    return dict(positions=positions, swaps=swaps, cut_moves=cut_moves, questions=len(ref), questions_exact=exact,
                max_swap_gap=max_gap)


def report(name, stats):
    line = dict(case=name, **stats)
    print("ranking %s: %d of %d questions identical, %d of %d positions swapped inside near-ties "
          "(max reference gap %.1e), %d cut moves" % (name, stats["questions_exact"], stats["questions"],
                                                      stats["swaps"], stats["positions"], stats["max_swap_gap"],
                                                      stats["cut_moves"]))
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = os.path.join(root, "gpurun_out")
    if os.path.isdir(out):
        with open(os.path.join(out, "ranking_swaps.jsonl"), "a") as f:
            f.write(json.dumps(line) + "\n")
    return line
