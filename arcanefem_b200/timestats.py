"""`time_stats.json` in the shape the reference's testlab module dumps (modules/testlab/FemModule.cc:19-56: cacheWarming,
nbParallelInstance, acceleratorRuntime, meshDim, nbNode, nbBoundaryElement, nbElement + the timer tree of
ITimeStats::dumpStatsJSON), so that the reference's own post-processing (modules/testlab/benchmarking/get_stats_from_json.py:
find_key("AssembleBilinearOperator_<Format>") -> {"Cumulative": "<seconds> ..."} with the sub-actions BuildMatrix /
AddAndCompute, divided by cacheWarming - 1) keeps working on numbers measured here."""
from __future__ import annotations

import json

# matrix-format option of the reference -> timer name of its assembly (modules/testlab/*BiliAssembly.cc: Timer::Action names)
TIMER_OF_FORMAT = {"legacy": "Legacy", "coo": "Coo", "coo-sorting": "CooSort", "coo-gpu": "Coo_Gpu", "coo-sorting-gpu": "CooSort_Gpu",
                   "csr": "Csr", "csr-gpu": "Csr_Gpu", "nwcsr": "CsrNodeWise", "blcsr": "CsrBuildLess"}


def _action(seconds, children=None):
    node = {"Cumulative": f"{seconds:.9g} (s)", "Local": f"{seconds - sum(c for c in (children or {}).values()):.9g} (s)"}
    for name, sec in (children or {}).items():
        node[name] = {"Cumulative": f"{sec:.9g} (s)", "Local": f"{sec:.9g} (s)"}
    return node


def time_stats(timed_steps, nb_parallel_instance, mesh_dim, nb_node, nb_boundary_element, nb_element, formats, accelerator_runtime="cuda"):
    """formats: {reference format option: (build_matrix_seconds_per_step, add_and_compute_seconds_per_step)}.
    The reference accumulates over cacheWarming iterations and the script divides by cacheWarming - 1 (the first one is
    the warm-up): cumulative times here cover `timed_steps` steps and cacheWarming = timed_steps + 1."""
    timers = {}
    for fmt, (bm, ac) in formats.items():
        name = "AssembleBilinearOperator_" + TIMER_OF_FORMAT.get(fmt, fmt)
        timers[name] = _action((bm + ac) * timed_steps, {"BuildMatrix": bm * timed_steps, "AddAndCompute": ac * timed_steps})
    return {"cacheWarming": timed_steps + 1, "nbParallelInstance": nb_parallel_instance, "acceleratorRuntime": accelerator_runtime, "meshDim": mesh_dim,
            "nbNode": nb_node, "nbBoundaryElement": nb_boundary_element, "nbElement": nb_element, "Timer": {"Main": timers}}


def write_time_stats(path, *a, **k):
    with open(path, "w") as f:
        json.dump(time_stats(*a, **k), f)
