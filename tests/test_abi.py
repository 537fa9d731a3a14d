"""CPU: the C-ABI library loads without a GPU and exports exactly the symbols include/qlidar.h declares."""
import os
import re
import subprocess

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "qlidar.h")


def _declared():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ql_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_exported_and_bound():
    from qlidar import _lib
    declared = _declared()
    assert len(declared) >= 18
    lib = _lib.lib()                       # raises if the .so is missing: there is no fallback
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in qlidar.h but not exported"
    assert sorted(_lib.SIGNATURES) == declared, "ctypes table and header disagree"
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r"\bT (ql_[a-z0-9_]+)", out))
    assert set(declared) <= exported


def test_host_only_entry_points():
    from qlidar import _lib
    lib = _lib.lib()
    assert lib.ql_abi_version() == 3
    assert lib.ql_error_string(0) == b"ok"
    assert b"workspace" in lib.ql_error_string(-4)
    assert lib.ql_hash_capacity(1000) == 2048
    assert lib.ql_hash_capacity(600000) == 2 ** 21
    assert lib.ql_rulebook_num_tiles(129) == 2
    assert lib.ql_rulebook_mask_words(27) == 1 and lib.ql_rulebook_mask_words(125) == 4
    # chunk = one kernel offset x one <=128-byte row segment, padded to 32 / 64 / 128 bytes
    assert lib.ql_packed_weight_bytes(16, 16, 27, _lib.QL_F16) == 27 * 16 * 32
    assert lib.ql_packed_weight_bytes(16, 16, 27, _lib.QL_S8) == 27 * 16 * 32       # 16-byte rows are zero padded to the 32-byte k-step
    assert lib.ql_packed_weight_bytes(128, 64, 27, _lib.QL_F16) == 27 * 2 * 64 * 128
    assert lib.ql_packed_weight_bytes(15, 16, 27, _lib.QL_F32) == 0
    # weights are shared-memory resident up to C = 32 (fp16) / C = 64 (int8) for a 3^3 kernel, streamed per unit above
    assert lib.ql_spconv_weights_streamed(32, 32, 27, _lib.QL_F16) == 0 and lib.ql_spconv_weights_streamed(64, 64, 27, _lib.QL_F16) == 1
    assert lib.ql_spconv_weights_streamed(64, 64, 27, _lib.QL_S8) == 0 and lib.ql_spconv_weights_streamed(128, 128, 27, _lib.QL_S8) == 1
    assert lib.ql_spconv_weights_streamed(128, 128, 3, _lib.QL_F16) == 0           # conv_out: 3 offsets fit
    assert lib.ql_rulebook_group_workspace_bytes(1000) >= 2000 + 512 * 4


def _chunk_geom(row_bytes):
    if row_bytes > 128:
        return 128, (row_bytes + 127) // 128
    return (32 if row_bytes <= 32 else 64 if row_bytes <= 64 else 128), 1


def _k_word_src(ch, c):
    """K order inside a chunk row: identity for CH = 32 (lane-per-row gather); for CH >= 64 the order in which the
    quad gather + tcgen05.st.16x256b leaves a row segment in tensor memory: column 8*v2 + 2*t0 + e <- word 2*(CH/32)*t0 + 2*v2 + e."""
    if ch < 64:
        return c
    v2, t0, e = c >> 3, (c >> 1) & 3, c & 1
    return 2 * (ch // 32) * t0 + 2 * v2 + e


def test_pack_weights_host_layout():
    """The packed image is, per (kernel offset, 128-byte row segment) chunk, the K-major swizzled [c_out x CH] shared-memory
    layout the tcgen05 B descriptors assume: 16-byte piece c of row r lives at
    (r//8)*8*CH + (r%8)*CH + ((c ^ x(r))*16), x(r) = r%8 (SWIZZLE_128B), (r//2)%4 (64B), (r//4)%2 (32B); the 4-byte K words
    of a row are in the gather's TMEM column order (_k_word_src)."""
    from qlidar import ops
    rng = np.random.default_rng(0)
    for dtype, cin, cout, K in [(torch.int8, 16, 16, 27), (torch.int8, 64, 32, 27), (torch.float16, 16, 48, 27),
                                (torch.float16, 128, 64, 3), (torch.float16, 32, 256, 125), (torch.int8, 128, 128, 27),
                                (torch.float16, 24, 16, 5)]:
        if dtype == torch.int8:
            w = torch.from_numpy(rng.integers(-127, 128, size=(cout, K, cin)).astype(np.int8))
        else:
            w = torch.from_numpy(rng.integers(-127, 128, size=(cout, K, cin)).astype(np.float32)).half()
        packed = ops.pack_weights(w).numpy()
        raw = w.contiguous().view(torch.uint8).reshape(cout, K, -1).numpy()
        row_bytes = raw.shape[2]
        ch, nseg = _chunk_geom(row_bytes)
        assert packed.size == K * nseg * cout * ch
        seen = np.zeros(packed.size, dtype=bool)
        for k in range(K):
            for seg in range(nseg):
                img = packed[(k * nseg + seg) * cout * ch:(k * nseg + seg + 1) * cout * ch]
                base = (k * nseg + seg) * cout * ch
                for r in range(cout):
                    x = r % 8 if ch == 128 else ((r // 2) % 4 if ch == 64 else (r // 4) % 2)
                    for c in range(ch // 4):                     # 4-byte K word c of the chunk row == TMEM column c of A
                        off = (r // 8) * 8 * ch + (r % 8) * ch + (((c // 4) ^ x) * 16) + 4 * (c % 4)
                        b0 = seg * 128 + 4 * _k_word_src(ch, c)
                        want = raw[r, k, b0:b0 + 4] if b0 < row_bytes else np.zeros(4, np.uint8)
                        assert np.array_equal(img[off:off + 4], want)
                        seen[base + off:base + off + 4] = True
        assert seen.all()


def test_ops_reject_cpu_tensors():
    import pytest
    from qlidar import ops
    from qlidar._lib import QlidarError
    with pytest.raises(QlidarError):
        ops.hash_build(torch.zeros((4, 4), dtype=torch.int32), None, (1, 2, 2, 2))


def test_spconv_import_alias_surface():
    """`import qlidar as spconv` satisfies the attribute paths the reference touches
    (pcdet/utils/spconv_utils.py:3-10,23; quant/quant.py:1-3; spconv_backbone.py:12-17)."""
    import qlidar as spconv
    assert float(spconv.__version__[2:]) >= 2.2
    spconv.constants.SPCONV_USE_DIRECT_TABLE = False
    sp = spconv.pytorch
    for name in ("SparseConvTensor", "SparseSequential", "SparseModule", "SubMConv3d", "SparseConv3d", "SubMConv2d",
                 "SparseConv2d", "SparseInverseConv3d", "replace_feature"):
        assert hasattr(sp, name), name
    assert issubclass(sp.SubMConv3d, spconv.conv.SparseConvolution)
    assert spconv.pytorch.modules.SparseModule is spconv.SparseModule
    for name in ("QConvNd", "QConv3d", "QConv2d", "GQConv3d", "q_conv3d", "collect_stats", "compute_amax", "TensorQuantizer",
                 "QuantDescriptor", "MeanVFE", "DynamicMeanVFE", "HeightCompression", "VoxelBackBone8x", "VoxelResBackBone8x",
                 "VoxelResBackBone8xVoxelNeXt", "VoxelGeneratorWrapper", "BackboneEngine", "shard"):
        assert hasattr(spconv, name), name
    conv = sp.SubMConv3d(16, 32, 3, padding=1, bias=False, indice_key="subm1")
    assert tuple(conv.weight.shape) == (32, 3, 3, 3, 16)            # spconv-2 layout (oc, kd, kh, kw, ic), quant/quant.py:37-39
    q = spconv.QConvNd(conv, 8, 8, True)
    assert hasattr(q, "w_quant") and hasattr(q, "act_quant") and q.module is conv   # names end in _quant: quantize.py:178
