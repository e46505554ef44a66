"""The mbarrier protocol of the pipelined backward kernel (attn_bwd2_kernel) checked by discrete-event simulation
(tools/sim_bwd2_protocol.py): random interleavings of the four warp roles must neither deadlock nor touch an operand
that is not ready.  Runs in about a second on the CPU."""
import importlib.util
import os
import random

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _sim():
    spec = importlib.util.spec_from_file_location("sim_bwd2", os.path.join(ROOT, "tools", "sim_bwd2_protocol.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_backward_v2_protocol_has_no_deadlock_or_stale_operand():
    sim = _sim()
    for seed in range(80):
        rng = random.Random(7000 + seed)
        items = [rng.choice([0, 1, 1, 2, 3, 5, 8]) for _ in range(rng.randint(1, 7))]
        sim.Sim(items, seed).run()


def test_the_simulator_detects_a_wrong_parity():
    sim = _sim()

    class Broken(sim.Sim):
        def producer(self):                      # waits for the wrong phase of res_empty
            for k, nvis in enumerate(self.items):
                yield ("wait", "res_empty", k & 1)
                self.later("res_full")

    try:
        Broken([2, 1], 0).run()
    except (RuntimeError, AssertionError):
        return
    raise AssertionError("a producer waiting on the wrong res_empty parity was not detected")
