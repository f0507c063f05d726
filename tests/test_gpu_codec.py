"""voxblox_msgs/Layer block codec on the device: cg_layer_serialize /
cg_layer_deserialize against a numpy restatement of voxblox Block::serializeToIntegers
(three words per voxel: float bits of distance, float bits of weight, colour
a | b<<8 | g<<16 | r<<24).  Reference call sites: coxgraph/include/coxgraph/utils/
msg_converter.h:49-50,107-109, coxgraph/src/client/map_server.cpp:88-89, tsdf_recover.h:95."""
import numpy as np
import pytest

from tests import util

pytestmark = pytest.mark.gpu


def serialize_to_integers(vox):
    """numpy statement of Block<TsdfVoxel>::serializeToIntegers for [B, 4096] voxels."""
    out = np.zeros((len(vox), 4096, 3), np.uint32)
    out[..., 0] = vox["distance"].view(np.uint32)
    out[..., 1] = vox["weight"].view(np.uint32)
    c = vox["rgba"].astype(np.uint32)
    out[..., 2] = c[..., 3] | (c[..., 2] << 8) | (c[..., 1] << 16) | (c[..., 0] << 24)
    return out.reshape(len(vox), 4096 * 3)


def test_serialize_matches_voxblox_block_format_and_round_trips(gpu_ctx):
    from coxgraph_b200 import Layer, TsdfIntegrator
    from oracle import oracle_py as orc
    ocfg, gcfg = util.make_cfgs()
    frames = util.small_frames(3, stride=8)
    ol, gl = orc.Layer(0.05), Layer(gpu_ctx, 0.05, max_blocks=2048)
    integ = TsdfIntegrator(gcfg, gl)
    for (T, p, c) in frames[:2]:
        ol.integrate(ocfg, T, p, c)
        integ.integratePointCloud(T, p, c)
    idx, data = gl.serializeLayerAsMsg()
    gi, gv, _ = gl.download()
    assert np.array_equal(idx, gi) and data.dtype == np.uint32 and data.shape == (len(gi), 12288)
    assert np.array_equal(data, serialize_to_integers(gv)), "device codec != serializeToIntegers"
    # the oracle's layer in the same wire format agrees within the parity tolerance
    oi, ov, _ = ol.download()
    assert np.array_equal(idx, oi)
    assert np.array_equal(data.reshape(-1, 4096, 3)[..., 1].view(np.float32) > 0, ov["weight"] > 0)
    # round trip through the wire format is lossless
    back = Layer(gpu_ctx, 0.05, max_blocks=2048)
    back.deserializeMsgToLayer(idx, data)
    bi, bv, bf = back.download()
    assert np.array_equal(bi, gi) and np.array_equal(bv.tobytes(), gv.tobytes())
    assert (bf & 1).all(), "deserialised blocks carry has_data"
    # only_updated: after a reset only the blocks touched by the next frame are sent
    gl.resetUpdated()
    idx0, _ = gl.serializeLayerAsMsg(only_updated=True)
    assert len(idx0) == 0
    T, p, c = frames[2]
    st = integ.integratePointCloud(T, p, c)
    idx1, data1 = gl.serializeLayerAsMsg(only_updated=True)
    assert len(idx1) == st.blocks_touched
    gi2, gv2, gf2 = gl.download()
    keep = (gf2 & 2) != 0
    assert np.array_equal(idx1, gi2[keep]) and np.array_equal(data1, serialize_to_integers(gv2[keep]))
    # cg_layer_download_updated: the same blocks as voxblox TsdfVoxel records (what an adapter that
    # mirrors the layer on the host copies back after integratePointCloud)
    ui, uv, uf = gl.download_updated()
    assert np.array_equal(ui, gi2[keep]) and np.array_equal(uv.tobytes(), gv2[keep].tobytes())
    assert np.array_equal(uf, gf2[keep]) and (uf & 2).all()
    gl.resetUpdated()
    assert len(gl.download_updated()[0]) == 0
    for L in (gl, back):
        L.close()
