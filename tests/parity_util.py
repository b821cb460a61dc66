"""Shared helpers for the parity tests (test infrastructure; may import oracle/)."""
import contextlib
import json
import os

import torch

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REPORT_PATH = os.path.join(_ROOT, "gpurun_out", "parity_report.jsonl")


def report(test, **metrics):
    """Append one line of measured parity numbers (worst error ratios, not just a pass bit) to
    gpurun_out/parity_report.jsonl; tools/summarise_parity.py turns the file into profiles/rNN_parity.md."""
    line = json.dumps(dict(test=test, **{k: (float(v) if isinstance(v, (int, float)) else v) for k, v in metrics.items()}))
    print("PARITY " + line)
    try:
        os.makedirs(os.path.dirname(REPORT_PATH), exist_ok=True)
        with open(REPORT_PATH, "a") as f:
            f.write(line + "\n")
    except OSError:
        pass


@contextlib.contextmanager
def conv_impl(name):
    """Selects the k3 convolution kernels for the enclosed CUDA calls: 'auto' (size heuristics, what a user gets),
    'umma' (tcgen05 forward / data gradient / weight gradient wherever structurally possible, default 3xFP16 operands),
    'umma_tf32' (the same kernels with 3xTF32 operands), 'ffma', 'direct'."""
    from deepatlas_b200 import _lib
    _lib.call("da_set_conv_impl", CONV_IMPLS[name][0])
    _lib.call("da_set_conv_split", CONV_IMPLS[name][1])
    try:
        yield
    finally:
        _lib.call("da_set_conv_impl", 0)
        _lib.call("da_set_conv_split", 0)


# Width of the band around zero (relative to max|pre-activation| of the layer) inside which the CUDA path and the oracle
# may take different activation branches: every kernel family (exact FFMA, 3xTF32, the default 3xFP16 on scaled
# operands) differs from ATen by fp32-level round-off only (2e-5 after a dozen layers with batch-1 BatchNorm).
MASK_BAND = {"auto": 2e-5, "umma": 2e-5, "umma_tf32": 2e-5, "ffma": 2e-5, "direct": 2e-5, "umma_f16x1": 5e-3}


def max_flips(replay):
    """Allowed number of activation-mask flips: a fixed floor plus the share of elements expected inside the band."""
    return 64 + int(4 * replay.eps * replay.count)


CONV_IMPLS = {"auto": (0, 0), "direct": (1, 0), "ffma": (2, 0), "umma": (3, 0), "umma_tf32": (3, 1), "umma_f16x1": (3, 2)}


def rel_err(a, b):
    """max-norm relative error ||a-b||_inf / max(||b||_inf, tiny)."""
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / max(float(b.abs().max()), 1e-30))


def cpu_state(module, dtype=torch.float32):
    return {k: (v.detach().cpu().to(dtype) if v.is_floating_point() else v.detach().cpu().clone())
            for k, v in module.state_dict().items()}


def oracle_joint_loss(model, batch, P, dtype=torch.float32, replay=None):
    """The joint step of deepatlas_b200/joint.py restated on the CPU with oracle/ref_port.py.  Returns
    (loss, {param_name: grad}) with parameter names as in JointModel ('seg.*', 'reg.*')."""
    batch = [t.detach().cpu() for t in batch]
    seg_sd = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running" not in k else v)
              for k, v in cpu_state(model.seg, dtype).items()}
    reg_sd = {k: v.clone().requires_grad_(True) for k, v in cpu_state(model.reg, dtype).items()}
    if replay is not None:
        replay.restart()
        with replay:
            loss = P.joint_loss(seg_sd, reg_sd, batch, model.n_classes, model.lambdas, dtype)
    else:
        loss = P.joint_loss(seg_sd, reg_sd, batch, model.n_classes, model.lambdas, dtype)
    loss.backward()
    grads = {}
    for k, v in seg_sd.items():
        if v.is_floating_point() and v.requires_grad and v.grad is not None:
            grads["seg." + k] = v.grad
    for k, v in reg_sd.items():
        if v.grad is not None:
            grads["reg." + k] = v.grad
    return loss.detach(), grads


def rel_err_quantile(a, b, frac=1e-4):
    """Like rel_err but ignoring the worst `frac` of the elements: used where an activation mask can flip on
    values within fp32 round-off of zero (a discontinuity of the reference function itself)."""
    a, b = a.detach().double().cpu().flatten(), b.detach().double().cpu().flatten()
    d = (a - b).abs()
    k = max(1, int(d.numel() * (1.0 - frac)))
    return float(d.kthvalue(k).values / max(float(b.abs().max()), 1e-30))


def check_grads_vs_truth(ours, ref32, truth64, tol, slack=3.0, floor=1e-3, strict=True):
    """Gradient parity on the precision ladder: fp64 oracle = truth, fp32 oracle = the
    reference's own behaviour.  Each of OUR gradients must be within max(tol, slack * reference-fp32 error) of
    the truth, errors normalised by max(||truth_k||_inf, floor * max_k ||truth_k||_inf) so that analytically-zero
    gradients (conv bias in front of a BatchNorm) are judged on an absolute scale.  Returns the worst ratio."""
    gmax = max(float(v.abs().max()) for v in truth64.values())
    worst = (0.0, None)
    worst_abs = (0.0, None)
    for k, t in truth64.items():
        assert k in ours, f"missing gradient {k}"
        scale = max(float(t.abs().max()), floor * gmax)
        e_ours = float((ours[k].detach().double().cpu() - t).abs().max()) / scale
        e_ref = float((ref32[k].detach().double() - t).abs().max()) / scale
        bound = max(tol, slack * e_ref)
        if strict:   # bench.py reports the worst ratio instead of stopping
            assert e_ours <= bound, f"grad {k}: ours-vs-fp64 {e_ours:.3e} > bound {bound:.3e} (reference fp32-vs-fp64 {e_ref:.3e})"
        if e_ours / bound > worst[0]:
            worst = (e_ours / bound, k)
        if e_ours > worst_abs[0]:
            worst_abs = (e_ours, k, e_ref)
    return {"worst_ratio_to_bound": worst[0], "worst_ratio_param": worst[1], "worst_rel_err_vs_fp64": worst_abs[0],
            "worst_rel_err_param": worst_abs[1], "reference_fp32_rel_err_same_param": worst_abs[2] if len(worst_abs) > 2 else None}


# ------------------------------------------------------------------------------------------------------------
# Activation-mask replay.  ReLU / LeakyReLU are piecewise linear: a pre-activation within fp32 round-off of zero
# may land on either side of the kink depending on summation order, and ONE such flip changes every weight
# gradient upstream by O(1e-2) relative (measured: tools/debug_layer_grads.py) although both answers are valid
# sub-gradients of the same function.  The whole-network parity tests therefore record the sign pattern of every
# activation output of the CUDA forward pass and make the oracle take the same branches; a mismatch is accepted
# only where the oracle's own pre-activation is within round-off of zero (asserted), so a real error cannot hide.
# ------------------------------------------------------------------------------------------------------------
class MaskRecorder:
    """Context manager: records (activation output > 0) of every activation the CUDA path applies, in call order."""

    def __init__(self):
        self.masks = []

    def __enter__(self):
        from deepatlas_b200 import networks, ops
        self._ops, self._networks = ops, networks
        self._bn_act, self._conv3d, self._leaky = ops.bn_act, ops.conv3d, networks._LeakyFunction.apply

        def bn_act(*a, **k):
            y = self._bn_act(*a, **k)
            slope = k.get("slope", a[8] if len(a) > 8 else None)
            if slope is not None:
                self.masks.append((y.detach() > 0).cpu())
            return y

        def conv3d(*a, **k):
            y = self._conv3d(*a, **k)
            if k.get("slope", None) is not None:
                self.masks.append((y.detach() > 0).cpu())
            return y

        def leaky(y_in, slope):
            y = self._leaky(y_in, slope)
            self.masks.append((y.detach() > 0).cpu())
            return y

        ops.bn_act, ops.conv3d, networks._LeakyFunction.apply = bn_act, conv3d, leaky
        return self

    def __exit__(self, *exc):
        self._ops.bn_act, self._ops.conv3d, self._networks._LeakyFunction.apply = self._bn_act, self._conv3d, self._leaky
        return False


class MaskReplay:
    """Context manager: oracle/ref_port.py activations take the recorded branches (see above).  ``flips`` counts the
    elements whose oracle sign differed; each must have |pre-activation| <= eps * max|pre-activation|."""

    def __init__(self, masks, eps=2e-5):
        self.masks, self.eps, self.flips, self.count, self._i = masks, eps, 0, 0, 0

    def restart(self):
        self._i = 0

    def __enter__(self):
        from oracle import ref_port as P
        self._P, self._act = P, P._act

        def act(x, kind):
            m = self.masks[self._i]
            self._i += 1
            assert m.shape == x.shape, f"activation {self._i - 1}: recorded {tuple(m.shape)} vs oracle {tuple(x.shape)}"
            diff = m != (x.detach() > 0)
            n = int(diff.sum())
            self.count += m.numel()
            if n:
                worst = float(x.detach()[diff].abs().max()) / max(float(x.detach().abs().max()), 1e-30)
                assert worst <= self.eps, f"activation {self._i - 1}: mask differs at |z|/max|z| = {worst:.2e} (not round-off)"
                self.flips += n
            slope = {"ReLU": 0.0, "LeakyReLU": 0.01}[kind]
            return torch.where(m, x, x * slope)

        P._act = act
        return self

    def __exit__(self, *exc):
        self._P._act = self._act
        return False
