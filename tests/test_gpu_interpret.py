"""-m gpu: device-side routing statistics against the sums produced by executing the reference's
own accumulation statements (oracle/gen_golden_tail.py: route_stats_case)."""
import os

import pytest
import torch

from helpers import GOLD, max_rel

pytestmark = pytest.mark.gpu


def test_routing_stats_accumulate_like_evaluate_epoch():
    from multimodalrouting_b200.interpret import RoutingStatsAccumulator
    gold = torch.load(os.path.join(GOLD, "tail_route_stats.pt"), weights_only=False)
    acc = RoutingStatsAccumulator(gold["K"])
    for b in gold["batches"]:
        acc.update(b["rc_raw"].cuda(), b["rc_report"].cuda(), b["prim_acts"].cuda())
    res = acc.result()
    assert res["num_samples"] == gold["num_samples"]
    for k in ("rc_raw_sum", "rc_report_sum", "eff_sum", "prim_act_sum", "avg_rc_report"):
        assert max_rel(res[k], gold[k]) < 2e-6, k
    # graph-capturable (no synchronisation, no allocation with static inputs) and resettable
    acc.reset()
    b = {k: v.cuda() for k, v in gold["batches"][0].items()}
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        acc.update(b["rc_raw"], b["rc_report"], b["prim_acts"])
    torch.cuda.current_stream().wait_stream(s)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        acc.update(b["rc_raw"], b["rc_report"], b["prim_acts"])
    g.replay()
    torch.cuda.synchronize()
    res = acc.result()
    assert res["num_samples"] == 2 * b["rc_raw"].shape[0]          # warm-up call + one replay (the capture itself runs nothing)
    assert max_rel(res["rc_raw_sum"], 2 * b["rc_raw"].float().sum(0).cpu()) < 2e-6


def test_routing_stats_rejects_bad_shapes():
    from multimodalrouting_b200.interpret import RoutingStatsAccumulator
    acc = RoutingStatsAccumulator(3)
    with pytest.raises(ValueError):
        acc.update(torch.zeros(4, 10, 4, device="cuda"), None, torch.zeros(4, 10, device="cuda"))
    with pytest.raises(ValueError):
        acc.update(torch.zeros(4, 10, 3, device="cuda"), None, torch.zeros(5, 10, device="cuda"))
