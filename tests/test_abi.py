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
    assert lib.ql_abi_version() == 1
    assert lib.ql_error_string(0) == b"ok"
    assert b"workspace" in lib.ql_error_string(-4)
    assert lib.ql_hash_capacity(1000) == 2048
    assert lib.ql_hash_capacity(600000) == 2 ** 21
    assert lib.ql_rulebook_num_tiles(129) == 2
    assert lib.ql_packed_weight_bytes(16, 16, 27, _lib.QL_F16) == 7 * 16 * 128
    assert lib.ql_packed_weight_bytes(16, 16, 27, _lib.QL_S8) == 4 * 16 * 128
    assert lib.ql_packed_weight_bytes(15, 16, 27, _lib.QL_F32) == 0


def test_pack_weights_host_layout():
    """The packed image is the K-major SWIZZLE_128B shared-memory layout the tcgen05 descriptors assume:
    byte (row r, 16B chunk c) of a 128-byte K stage lives at (r//8)*1024 + (r%8)*128 + ((c ^ (r%8))*16)."""
    from qlidar import ops
    rng = np.random.default_rng(0)
    for dtype, cin, cout, K in [(torch.int8, 16, 16, 27), (torch.int8, 64, 32, 27), (torch.float16, 16, 48, 27),
                                (torch.float16, 128, 64, 3), (torch.float16, 32, 256, 125)]:
        if dtype == torch.int8:
            w = torch.from_numpy(rng.integers(-127, 128, size=(cout, K, cin)).astype(np.int8))
        else:
            w = torch.from_numpy(rng.integers(-127, 128, size=(cout, K, cin)).astype(np.float32)).half()
        packed = ops.pack_weights(w).numpy()
        raw = w.contiguous().view(torch.uint8).reshape(cout, -1).numpy()
        kbytes = raw.shape[1]
        stages = (kbytes + 127) // 128
        assert packed.size == stages * cout * 128
        flat = np.zeros((cout, stages * 128), dtype=np.uint8)
        flat[:, :kbytes] = raw
        for s in range(stages):
            img = packed[s * cout * 128:(s + 1) * cout * 128]
            for r in range(cout):
                for c in range(8):
                    off = (r // 8) * 1024 + (r % 8) * 128 + ((c ^ (r % 8)) * 16)
                    assert np.array_equal(img[off:off + 16], flat[r, s * 128 + c * 16:s * 128 + c * 16 + 16])


def test_ops_reject_cpu_tensors():
    import pytest
    from qlidar import ops
    from qlidar._lib import QlidarError
    with pytest.raises(QlidarError):
        ops.hash_build(torch.zeros((4, 4), dtype=torch.int32), None, (1, 2, 2, 2))
