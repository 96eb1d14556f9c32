"""The fused kick-drift-kick kernels through the C ABI: vectorised (16-byte aligned) and element-wise paths must be
bit-identical to torch's separately rounded `v + a*(dt/2)`, `x + v*dt` (simulation.py:132-141), for every phase,
dtype, dimension and ragged size; the packed records they emit must equal nb_pack_sources of the new positions."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _case(n, dim, dtype, offset):
    g = torch.Generator().manual_seed(n * 7 + dim)
    big = [torch.randn(n + 3, dim, generator=g).to(dtype).to(DEV) for _ in range(3)]
    x, v, a = (b[offset:offset + n] for b in big)          # offset != 0 -> base pointers not 16-byte aligned
    m = (torch.rand(n + 3, generator=g) + 0.5).to(dtype).to(DEV)[offset:offset + n]
    return x, v, a, m


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
@pytest.mark.parametrize("dim", [2, 3])
@pytest.mark.parametrize("n", [1, 2, 3, 4, 5, 255, 256, 257, 1000])
@pytest.mark.parametrize("offset", [0, 1])
def test_kdk_phases_bit_exact(n, dim, dtype, offset):
    from nbody_cosmological_simulation_b200 import _lib as L
    from nbody_cosmological_simulation_b200.ops import CudaOps
    ops = CudaOps()
    lib = ops.lib
    x, v, a, m = _case(n, dim, dtype, offset)
    assert x.is_contiguous()
    dt = 0.0137
    scal = ops.new_scalars(x.device)
    code = L.dtype_code(x)
    half = dt / 2
    for phase in (L.KDK_KICK_DRIFT, L.KDK_KICK, L.KDK_KICK_KICK_DRIFT):
        packed = torch.zeros(lib.nb_packed_bytes(n, dim, code) + 2 * lib.nb_chunk_bytes(dim, code), dtype=torch.uint8, device=x.device)
        total_chunks = lib.nb_num_chunks(n, code) + 2                      # two whole padding chunks as well
        want_packed = torch.zeros_like(packed)
        xo, vo = ops.kdk(phase, x, v, a.clone(), m, dt, 0, scal, packed=packed if phase != L.KDK_KICK else None,
                         total_chunks=total_chunks if phase != L.KDK_KICK else 0)
        vw = v + a * half                                                  # torch: mul and add separately rounded
        if phase == L.KDK_KICK_KICK_DRIFT:
            vw = vw + a * half
        assert torch.equal(vo, vw)
        if phase == L.KDK_KICK:
            assert xo is None
            continue
        xw = x + vw * dt
        assert torch.equal(xo, xw)
        ops.pack(xw.contiguous(), m, want_packed, total_chunks)
        assert torch.equal(packed, want_packed)


def test_kdk_snap_matches_free_standing_grid_quantize():
    """snap_levels > 0: accelerations are snapped to the linear grid in the same pass (quantize_force, INT4)."""
    from nbody_cosmological_simulation_b200 import _lib as L, quantization as Q
    from nbody_cosmological_simulation_b200.ops import CudaOps
    ops = CudaOps()
    for n, offset in ((1000, 0), (1001, 1)):
        x, v, a, m = _case(n, 2, torch.float32, offset)
        scal = ops.new_scalars(x.device)
        lo, hi = a.min().item(), a.max().item()
        scal[L.SLOT_ACC_MIN] = ops.lib.nb_key_from_double(lo)
        scal[L.SLOT_ACC_MAX] = ops.lib.nb_key_from_double(hi)
        want_a = Q._grid_quantize(a.contiguous(), 16)
        a_work = a.clone()
        _, vo = ops.kdk(L.KDK_KICK, None, v, a_work, m, 0.01, 16, scal)
        assert torch.equal(a_work, want_a)                                  # written back snapped
        assert torch.equal(vo, v + want_a * 0.005)
        assert len(torch.unique(a_work)) <= 16
