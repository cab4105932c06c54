"""Time-to-solution reporting: R99 and its bootstrap, restated from the reference's plotting
library so the TTS half of the headline metric can be reported without matplotlib/json_stream
(``ccvmplotlib/utils/sampleTTSmetric.py:123-214``, ``problem_metadata/boxqp_metadata.py:83-137``).
Host-side numpy/scipy bookkeeping -- not part of the device hot path."""
import numpy as np

GAP_KEYS = ("optimal", "one_percent", "two_percent", "three_percent", "four_percent", "five_percent", "ten_percent")


def calc_r99(success_probability, confidence=0.99):
    """Independent runs needed to see a success at least once with ``confidence``
    (sampleTTSmetric.py:123-158): inf at p=0, 1 at p=1, max(1, ln(1-conf)/ln(1-p)) otherwise."""
    if not 0 < confidence < 1:
        raise ValueError("confidence must be between 0 and 1")
    if success_probability == 0:
        return np.inf
    if success_probability == 1:
        return 1.0
    r99 = np.log(1 - confidence) / np.log(1 - success_probability)
    return 1.0 if r99 < 1 else float(r99)


def r99_distribution(success_probabilities, num_repeats, percentile=50.0, num_bootstraps=100, seed=1,
                     confidence=0.99):
    """Bootstrap distribution of the ``percentile``-th R99 over a set of same-size problems
    (sampleTTSmetric.py:160-214): Beta(0.5 + successes, 0.5 + failures) posterior per problem,
    ``num_bootstraps`` resamples drawn through the posterior's inverse CDF."""
    from scipy.stats import beta as beta_distribution
    rng = np.random.RandomState(seed)
    probs = list(success_probabilities)
    posterior = [(0.5 + p * num_repeats, 0.5 + (1 - p) * num_repeats) for p in probs]
    out = np.empty(num_bootstraps, dtype=float)
    for i in range(num_bootstraps):
        idx = rng.randint(0, len(posterior), len(probs))
        cdf = rng.uniform(0, 1, len(probs))
        sampled = [calc_r99(float(beta_distribution.ppf(c, *posterior[j])), confidence) for j, c in zip(idx, cdf)]
        out[i] = np.percentile(sampled, percentile)
    return out


def time_to_solution(solve_times, success_probabilities, num_repeats, percentile=50.0, **kw):
    """TTS = mean(per-run machine time) x mean(bootstrapped R99 percentile)
    (boxqp_metadata.py:117-135 with machine_time = mean solve_time, ccvm_solver.py:368-390)."""
    dist = r99_distribution(success_probabilities, num_repeats, percentile, **kw)
    return float(np.mean(solve_times) * np.mean(dist))


def tts_table(metadata, gap="optimal", percentiles=(25.0, 50.0, 75.0), **kw):
    """{problem_size: {percentile: TTS}} from a list of ``Solution.get_metadata_dict()`` records."""
    by_size = {}
    for rec in metadata:
        by_size.setdefault(rec["problem_size"], []).append(rec)
    table = {}
    for n, recs in sorted(by_size.items()):
        probs = [r["solution_performance"][gap] for r in recs]
        times = [r["solve_time"] for r in recs]
        reps = recs[0]["batch_size"]
        table[n] = {pc: time_to_solution(times, probs, reps, pc, **kw) for pc in percentiles}
    return table
