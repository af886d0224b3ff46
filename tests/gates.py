"""Test helper: collect the branch decisions (ReLU / LeakyReLU gates, max-pool routing) the CUDA path took in one
step, keyed like oracle.terra_oracle.gate_tape expects, so the oracle can replay them (see the comment there)."""
import torch
import torch.nn.functional as F

from tg_b200 import plan as P


def _nchw(t):
    if t.dim() == 5:
        t = P.from_parity_split(t) if t.shape[1] == 4 else t[:, 0]
    return t.permute(0, 3, 1, 2)


def arm(G=None, D=None, crit=None):
    """Switch the engines' activation traces on (before the forward passes)."""
    if G is not None:
        G._trace = {}
    if D is not None:
        D._engine.trace = []
    if crit is not None:
        crit._vgg_engine.trace = []


def collect(G=None, D=None, crit=None) -> dict:
    gates = {}
    if G is not None:
        for k, t in G._trace.items():
            if k.endswith(".y"):
                gates[k[:-1]] = (_nchw(t) > 0).cpu()
    if D is not None:
        for p, tr in enumerate(D._engine.trace):
            for idx, t in tr.items():
                gates[f"D{p}.{idx}"] = (_nchw(t) > 0).cpu()
    if crit is not None:
        for p, tr in enumerate(crit._vgg_engine.trace):
            for idx, t in tr.items():
                y = _nchw(t).float()
                gates[f"vgg{p}.{idx}"] = (y > 0).cpu()
                if idx in (2, 7):
                    gates[f"vgg{p}.pool{idx}"] = F.max_pool2d(y, 2, 2, return_indices=True)[1].cpu()
        tr = crit._vgg_engine.trace
        if len(tr) >= 2:       # perceptual L1 (losses.py:86-89): sign(features(input) - features(target))
            gates["l1.perceptual"] = torch.sign(_nchw(tr[0][14]).float() - _nchw(tr[1][14]).float()).cpu()
    return gates


def disarm(G=None, D=None, crit=None):
    if G is not None:
        G._trace = None
    if D is not None:
        D._engine.trace = None
    if crit is not None:
        crit._vgg_engine.trace = None


def summarize(tape) -> str:
    flips = sum(v[0] for v in tape.report.values())
    total = sum(v[1] for v in tape.report.values())
    worst = max((v[2] for v in tape.report.values()), default=0.0)
    return f"{flips} of {total} branch decisions differ ({flips / max(total, 1):.1e}); largest |pre|/std among them {worst:.1e}"
