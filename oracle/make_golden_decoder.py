"""Generate tests/golden/decoder_*.npz from the REAL reference decoder (authoring container only).

    python -m oracle.make_golden_decoder [case ...]     # needs /root/reference/src (read-only import); no names = every case

``encoding.networks.AutoEncoderGroupSkip`` is imported unmodified and fed ``oracle.decoder_ref.synthetic_state_dict``
weights; ``decode_batch`` / ``decode_grid`` live in ``encoding/model.py``, which does not import here (tensorboardX,
open3d ... are absent), so their chunk loop + clamp is driven around the real ``net.decode`` exactly as
model.py:319-349 does, and ``sample_grid_points_aabb`` is executed from the reference source file (its function
definition is extracted with ``ast`` at run time; nothing is copied into this repository).
The oracle restatement is asserted against the reference on the spot.
"""
import ast
import contextlib
import io
import os
import sys

import numpy as np
import torch

REF = "/root/reference/src"
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")

from oracle import decoder_ref as de
from oracle.cases import DECODER_CASES, ENCODER_CASES, make_decoder_inputs, make_encoder_inputs


def ref_grid_fn():
    src = open(os.path.join(REF, "encoding", "utils3d.py")).read()
    fn = next(n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and n.name == "sample_grid_points_aabb")
    ns = {"torch": torch}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), "utils3d.py", "exec"), ns)
    return ns["sample_grid_points_aabb"]


def main():
    sys.path.insert(0, REF)
    from encoding.networks import AutoEncoderGroupPBR, AutoEncoderGroupSkip, AutoEncoderGroupV3
    grid_fn = ref_grid_fn()
    os.makedirs(OUT, exist_ok=True)
    only = set(sys.argv[1:])
    for name, case in DECODER_CASES.items():
        if only and name not in only:
            continue
        spec = de.DecoderSpec(**case["spec"])
        sd = de.synthetic_state_dict(spec, case["wseed"])
        cls = AutoEncoderGroupPBR if spec.net_kind == "pbr" else (AutoEncoderGroupV3 if spec.mlp_kind == "base" else AutoEncoderGroupSkip)
        with contextlib.redirect_stdout(io.StringIO()):
            net = cls(spec.geo_feat_channels, spec.tex_feat_channels, spec.feat_channel_up,
                      spec.mlp_hidden_channels, spec.mlp_hidden_layers, use_tex=spec.use_tex,
                      tex_channels=spec.tex_channels)
        assert [k for k, _ in de.param_shapes(spec)] == list(net.state_dict().keys()), "state_dict key order differs"
        net.load_state_dict(sd)
        net.eval()
        maps, pts, aabb = make_decoder_inputs(case)
        save = dict(xy=maps[0].numpy(), xz=maps[1].numpy(), yz=maps[2].numpy(), aabb=aabb.numpy())
        with torch.no_grad():
            if "grid" in case:
                coords = grid_fn(aabb, case["grid"])
                assert torch.equal(coords, de.grid_points(aabb, case["grid"])), "grid points differ"
                pts = coords.view(-1, 3)
                save["grid_shape"] = np.array(coords.shape[:3])
                bs = 1000       # several ragged chunks
            else:
                save["pts"] = pts.numpy()
                bs = 256
            # model.py:319-333
            preds = torch.cat([net.decode(pts[i:i + bs], maps, aabb=aabb) for i in range(0, pts.shape[0], bs)], dim=0)
            preds[..., 1:] = preds[..., 1:].clamp_(0, 1)
            # the up-convolved planes the points are sampled from (for localising a failing kernel)
            g = spec.geo_feat_channels
            fp = {"geo": net.geo_convs([fm[:, :g] for fm in maps])}
            if spec.use_tex:
                fp["tex"] = net.tex_convs([fm[:, g:] for fm in maps])
        got = de.decode_batch(sd, spec, maps, pts, batch_size=bs, aabb=aabb)
        err = (got - preds).abs().max().item()
        assert err <= 2e-5 * max(1.0, preds.abs().max().item()), (name, err)
        ofp = de.feature_planes(sd, spec, maps)
        for br in fp:
            for pl, a, b in zip(de.PLANES, fp[br], ofp[br]):
                assert (a - b).abs().max().item() <= 2e-5 * max(1.0, a.abs().max().item())
                save[f"planes/{br}/{pl}"] = a.numpy()
        save["out"] = preds.numpy()
        np.savez_compressed(os.path.join(OUT, f"decoder_{name}.npz"), **save)
        print("decoder", name, "n", pts.shape[0], "oracle-vs-ref max abs", err, "out absmax", preds.abs().max().item(),
              "sdf range", preds[:, 0].min().item(), preds[:, 0].max().item())

    # ---- encoder half (networks.py:164-180)
    for name, case in ENCODER_CASES.items():
        if only and ("enc_" + name) not in only:
            continue
        spec = de.DecoderSpec(**case["spec"])
        sd = de.synthetic_state_dict(spec, case["wseed"])
        with contextlib.redirect_stdout(io.StringIO()):
            net = (AutoEncoderGroupPBR if spec.net_kind == "pbr" else AutoEncoderGroupSkip)(spec.geo_feat_channels, spec.tex_feat_channels, spec.feat_channel_up,
                                       spec.mlp_hidden_channels, spec.mlp_hidden_layers, use_tex=spec.use_tex,
                                       tex_channels=spec.tex_channels)
        net.load_state_dict(sd)
        net.eval()
        vol = make_encoder_inputs(case)
        with torch.no_grad():
            want = net.encode(vol)
        got = de.encode(sd, spec, vol)
        err = max((a - b).abs().max().item() for a, b in zip(want, got))
        assert err <= 1e-6, (name, err)
        np.savez_compressed(os.path.join(OUT, f"encoder_{name}.npz"), xy=want[0].numpy(), xz=want[1].numpy(), yz=want[2].numpy())
        print("encoder", name, [tuple(w.shape) for w in want], "oracle-vs-ref max abs", err)


if __name__ == "__main__":
    main()
