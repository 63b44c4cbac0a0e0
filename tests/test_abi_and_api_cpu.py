"""CPU tier: the C-ABI library loads and exports every symbol include/chromobius_b200.h declares; the Python mirror
validates arguments like the reference; nothing computes without a GPU."""
import ctypes
import os
import pickle
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope='module')
def cuda_lib():
    import __graft_entry__ as g
    return ctypes.CDLL(g.build_cuda())


def test_every_declared_symbol_is_exported(cuda_lib):
    header = open(os.path.join(ROOT, 'include', 'chromobius_b200.h')).read()
    names = sorted(set(re.findall(r'\b(cb200_[a-z_0-9]+)\s*\(', header)))
    assert len(names) >= 14
    for n in names:
        assert hasattr(cuda_lib, n), n


def test_product_does_not_link_the_oracle():
    import subprocess
    import __graft_entry__ as g
    out = subprocess.run(['nm', '-D', '--defined-only', g.build_cuda()], capture_output=True, text=True).stdout
    assert 'orc_' not in out and 'cbsim_' not in out
    for root, _, files in os.walk(os.path.join(ROOT, 'chromobius_b200')):
        for f in files:
            if f.endswith(('.py', '.h', '.cc', '.cu')):
                text = open(os.path.join(root, f)).read()
                assert 'import oracle' not in text and 'oracle/' not in text.replace('the oracle', ''), f


def test_decode_kernel_keeps_its_state_out_of_local_memory():
    """The per-shot state (`Shot`, ~650 B of table views and arena pointers) lives in shared memory, one copy per warp.
    A per-thread copy in local memory is what the kernel had until round 2 -- twice the global-load sectors of the whole
    kernel -- and it comes back silently as soon as someone declares `Shot s;` in a kernel again. The cubin's resource
    usage shows it: the stack frame grows by the size of the object."""
    import shutil
    import subprocess
    import __graft_entry__ as g
    if not shutil.which('cuobjdump'):
        pytest.skip('cuobjdump not available')
    out = subprocess.run(['cuobjdump', '-res-usage', g.build_cuda()], capture_output=True, text=True).stdout
    rows = re.findall(r'Function (\S*k_decodeILi(\d+)ELb0\S*):\s*\n\s*REG:(\d+) STACK:(\d+) SHARED:(\d+)', out)
    assert len(rows) >= 5, out[:500]
    for name, group, reg, stack, shared in rows:
        assert int(stack) <= 512, (group, stack)
        assert int(shared) >= 2048, (group, shared)  # the warps' Shot objects (static shared memory)
        if int(group) == 256:
            assert int(reg) <= 80, reg  # three 256-thread groups per SM


def test_decode_kernel_solver_classes_are_the_measured_configuration():
    """Which blossom solver each size class runs is a measured choice: the run-time layout of components
    over 32 events is solved by the whole warp (solve_w; the master / worker command dispatch is gone from the device code),
    the fixed 16- and 32-vertex layouts by the sequential solver. The cubin's symbols show what was compiled in."""
    import shutil
    import subprocess
    import __graft_entry__ as g
    if not shutil.which('cuobjdump') or not shutil.which('c++filt'):
        pytest.skip('cuobjdump / c++filt not available')
    elf = subprocess.run(['cuobjdump', '-elf', g.build_cuda()], capture_output=True, text=True).stdout
    mangled = sorted(set(re.findall(r'_ZN5cb2007Blossom\w+', elf)))
    names = subprocess.run(['c++filt'], input='\n'.join(mangled), capture_output=True, text=True).stdout
    assert re.search(r'Blossom<.*DevGroup<256>, 0>::solve_w', names), names[:2000]
    assert 'worker_loop' not in names and '::dispatch(' not in names
    assert not re.search(r'Blossom<.*DevGroup<256>, (16|32)>::solve_w', names)
    assert re.search(r'Blossom<.*DevGroup<256>, 16>::matching', names) and re.search(r'Blossom<.*DevGroup<256>, 32>::matching', names)


def test_compile_without_gpu_fails_loudly(cuda_lib):
    import chromobius_b200 as cb
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip('a GPU is present')
    except ImportError:
        pass
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        cb._compile_text('error(0.1) D0 L0\ndetector(0,0,0,0) D0')


def test_invalid_dem_reports_the_reference_message(cuda_lib):
    import chromobius_b200 as cb
    with pytest.raises(ValueError, match='4 coordinates'):
        cb._compile_text('error(0.1) D0\ndetector D0')


def test_from_dem_type_check():
    # pybind/chromobius_pybind_test.py:36-41
    import chromobius_b200 as cb
    with pytest.raises(ValueError, match='must be a stim.DetectorErrorModel'):
        cb.compile_decoder_for_dem('error(0.1) D0')
    with pytest.raises(ValueError, match='must be a stim.DetectorErrorModel'):
        cb.CompiledDecoder.from_dem(object())


def test_version_and_sinter_surface():
    # pybind/chromobius_pybind_test.py:22-24, pybind/sinter_compat_pybind_test.py:18-21
    import chromobius_b200 as cb
    assert isinstance(cb.__version__, str) and cb.__version__.startswith('1.2')
    d = cb.sinter_decoders()
    assert list(d) == ['chromobius']
    dec = d['chromobius']
    assert hasattr(dec, 'compile_decoder_for_dem') and hasattr(dec, 'decode_via_files')
    clone = pickle.loads(pickle.dumps(dec))
    assert clone == dec and not (clone != dec)
    with pytest.raises(TypeError):
        dec.compile_decoder_for_dem('positional is not allowed')


class _FakeLib:
    """Stands in for the shared library so that argument validation can be tested without a GPU."""

    def __init__(self, nb, ob):
        self.nb, self.ob = nb, ob
        self.calls = 0

    def cb200_num_detectors(self, h): return self.nb * 8
    def cb200_num_observables(self, h): return self.ob * 8
    def cb200_num_detector_bytes(self, h): return self.nb
    def cb200_num_observable_bytes(self, h): return self.ob
    def cb200_free(self, h): pass

    def cb200_decode_batch(self, *a):
        self.calls += 1
        return 0


def test_predict_argument_validation():
    # pybind/chromobius.pybind.cc:49-93
    import chromobius_b200 as cb
    fake = _FakeLib(3, 1)
    dec = cb.CompiledDecoder(1, fake)
    with pytest.raises(ValueError, match="dets.dtype wasn't np.uint8"):
        dec.predict_obs_flips_from_dets_bit_packed(np.zeros((2, 3), np.bool_))
    with pytest.raises(ValueError, match='dets.shape not in'):
        dec.predict_obs_flips_from_dets_bit_packed(np.zeros((2, 3, 1), np.uint8))
    with pytest.raises(ValueError, match='num_detector_bytes=3'):
        dec.predict_obs_flips_from_dets_bit_packed(np.zeros((2, 4), np.uint8))
    with pytest.raises(ValueError, match="dets.stride\\[1\\] wasn't equal to 1"):
        dec.predict_obs_flips_from_dets_bit_packed(np.zeros((2, 6), np.uint8)[:, ::2])
    assert fake.calls == 0
    out = dec.predict_obs_flips_from_dets_bit_packed(np.zeros((5, 3), np.uint8))
    assert out.shape == (5, 1) and out.dtype == np.uint8
    out, w = dec.predict_weighted_obs_flips_from_dets_bit_packed(np.zeros((5, 8), np.uint8)[:, :3])  # padded rows allowed
    assert out.shape == (5, 1) and w.shape == (5,) and w.dtype == np.float32
    out, w = dec.predict_weighted_obs_flips_from_dets_bit_packed(np.zeros(3, np.uint8))  # 1-D: outputs lose the shot axis
    assert out.shape == (1,) and w.shape == ()
    assert fake.calls == 3


def test_models_the_reference_accepts_and_this_library_rejects_are_reported(built):
    """include/chromobius_b200.h documents them: probabilities above 0.5 and zero-weight edges (PyMatching handles both)."""
    from sim_util import SimDecoder
    coords = 'detector(0,0,0,0) D0\ndetector(0,0,0,1) D1\n'
    with pytest.raises(ValueError, match='above 0.5'):
        SimDecoder('error(0.6) D0 D1\n' + coords)
    with pytest.raises(ValueError, match='non-positive weight'):
        SimDecoder('error(0.5) D0 D1\n' + coords)
    SimDecoder('error(0.4) D0 D1\n' + coords)
