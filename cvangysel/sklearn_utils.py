"""reference: cvangysel-common/py/cvangysel/sklearn_utils.py:7-33 picked a sklearn NearestNeighbors
back-end for a metric.  The GPU ranker does not use sklearn; the function is kept so callers importing it
keep working and returns the exact tree search that 'auto' resolved to for Euclidean distance."""


def neighbors_algorithm(metric):
    return 'kd_tree' if metric in ('euclidean', 'l2', 'minkowski', 'manhattan', 'chebyshev') else 'brute'
