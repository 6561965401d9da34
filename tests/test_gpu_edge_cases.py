"""Adversarial small cases through the CUDA path: results and IndexError behaviour == the oracle's."""
import pytest
import torch

from conftest import assert_close_to_max
from edge_cases import cases

pytestmark = pytest.mark.gpu


def _run(fn):
    try:
        return fn(), None
    except (IndexError, ValueError) as e:
        return None, type(e)


@pytest.mark.parametrize("variant", [None, "global_red", "vector_red", "smem_cache"])
def test_cuda_matches_oracle_on_edge_cases(oracle, variant, capsys):
    import event_utils_b200 as eu
    from event_utils_b200.representations.image import events_to_image_torch
    from event_utils_b200.representations.voxel_grid import events_to_voxel_torch
    eu.config.variant = variant
    try:
        for c in cases(2024, 150):
            D = [torch.from_numpy(c[k]).cuda() for k in "xytp"]
            hw = (c["H"], c["W"])
            o, oe = _run(lambda: oracle.voxel_f32(c["x"], c["y"], c["t"], c["p"], c["B"], hw))
            g, ge = _run(lambda: events_to_voxel_torch(*D, c["B"], sensor_size=hw).cpu().numpy())
            assert (oe is None) == (ge is None), ("voxel error behaviour", c["k"], oe, ge)
            if oe is None:
                assert_close_to_max(g, o, 1e-5, "voxel case %d" % c["k"])
            for interp in (None, "bilinear"):
                kw = dict(sensor_size=hw, clip_out_of_range=c["clip"], interpolation=interp, padding=c["padding"])
                o, oe = _run(lambda: oracle.image_torch_f32(c["x"], c["y"], c["p"], **kw))
                g, ge = _run(lambda: events_to_image_torch(D[0], D[1], D[3], **kw).cpu().numpy())
                assert (oe is None) == (ge is None), ("image error behaviour", c["k"], interp, kw, oe, ge)
                if oe is None:
                    assert_close_to_max(g, o, 1e-5, "image case %d %s" % (c["k"], interp))
    finally:
        eu.config.variant = None
        capsys.readouterr()
