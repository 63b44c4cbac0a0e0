"""The shared decoder service for sinter workers (chromobius_b200/service.py).

CPU tier: protocol, caching and concurrency with an injected oracle-backed decoder (test infrastructure -- the product
default is the CUDA decoder). GPU tier: the same through the real decoder and the sinter adapter."""
import multiprocessing as mp
import os

import numpy as np
import pytest

from dem_util import read_dem, NumpyDemSampler


class _OracleBacked:
    """Stands in for chromobius_b200.CompiledDecoder in the CPU tier."""

    def __init__(self, text):
        import oracle
        self._o = oracle.OracleDecoder(text)
        self.num_detectors = self._o.num_detectors
        self.num_observables = self._o.num_observables

    def predict_obs_flips_from_dets_bit_packed(self, dets):
        obs, _ = self._o.decode_batch(np.ascontiguousarray(dets))
        nb = (self.num_observables + 7) // 8
        out = np.zeros((len(dets), nb), np.uint8)
        for k in range(nb):
            out[:, k] = (obs >> np.uint64(8 * k)).astype(np.uint8)
        return out


def _worker(address, name, seed, shots, q):
    os.environ['CHROMOBIUS_B200_SERVICE'] = address
    import chromobius_b200
    text = read_dem(name)
    dets, _ = NumpyDemSampler(text).sample(shots, seed)
    dec = chromobius_b200.sinter_decoders()['chromobius'].compile_decoder_for_dem(dem=text)
    pred = dec.decode_shots_bit_packed(bit_packed_detection_event_data=dets)
    q.put((seed, pred.tobytes(), pred.shape))


def _run_workers(address, name, seeds, shots):
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(address, name, s, shots, q)) for s in seeds]
    for p in procs:
        p.start()
    out = {}
    for _ in procs:
        seed, raw, shape = q.get(timeout=300)
        out[seed] = np.frombuffer(raw, np.uint8).reshape(shape)
    for p in procs:
        p.join(timeout=60)
    return out


def test_service_protocol_with_injected_decoder(built, tmp_path):
    from chromobius_b200 import service
    address = str(tmp_path / 'svc.sock')
    svc = service.DecoderService(address, compile_fn=_OracleBacked)
    thread = svc.start_in_thread()
    name = 'midout_color_code_d5_r10_p1000'
    text = read_dem(name)
    try:
        got = _run_workers(address, name, seeds=[1, 2, 3], shots=300)
        ref = _OracleBacked(text)
        for seed, pred in got.items():
            dets, _ = NumpyDemSampler(text).sample(300, seed)
            assert (pred == ref.predict_obs_flips_from_dets_bit_packed(dets)).all()
        client = service.RemoteCompiledDecoder(address, text)
        st = client.service_stats()
        assert st['compiles'] == 1 and st['compile_hits'] == 3 and st['decode_calls'] == 3 and st['shots'] == 900
        with pytest.raises(ValueError):
            client.predict_obs_flips_from_dets_bit_packed(np.zeros((2, client.num_detector_bytes + 1), np.uint8))
        with pytest.raises(ValueError):   # compile errors come back with the reference's message
            service.RemoteCompiledDecoder(address, 'error(0.1) D0 L0\n')
        client.shutdown_service()
    finally:
        svc._stop.set()
        thread.join(timeout=5)


@pytest.mark.gpu
def test_service_on_gpu_matches_local_decoder(built, tmp_path):
    import chromobius_b200
    from chromobius_b200 import service
    address = str(tmp_path / 'svc.sock')
    svc = service.DecoderService(address, device=0)
    thread = svc.start_in_thread()
    name = 'midout_color_code_d9_r36_p1000'
    text = read_dem(name)
    try:
        got = _run_workers(address, name, seeds=[5, 6], shots=4000)
        local = chromobius_b200._compile_text(text, device=0)
        for seed, pred in got.items():
            dets, _ = NumpyDemSampler(text).sample(4000, seed)
            assert (pred == local.predict_obs_flips_from_dets_bit_packed(dets)).all()
        client = service.RemoteCompiledDecoder(address, text)
        assert client.service_stats()['compiles'] == 1
        client.shutdown_service()
    finally:
        svc._stop.set()
        thread.join(timeout=5)


def test_service_refuses_unsafe_socket_locations(tmp_path):
    """The socket may only live in a directory other users cannot write, and a non-socket at the address is never removed
    (the wire format itself is raw length-prefixed frames: nothing a peer sends is unpickled or evaluated)."""
    import stat
    from chromobius_b200 import service
    shared = tmp_path / 'shared'
    shared.mkdir()
    os.chmod(shared, 0o777)
    with pytest.raises(PermissionError):
        service.DecoderService(str(shared / 'svc.sock'), compile_fn=_OracleBacked)
    private = tmp_path / 'private'
    private.mkdir(mode=0o700)
    victim = private / 'data.txt'
    victim.write_text('keep me')
    with pytest.raises(FileExistsError):
        service.DecoderService(str(victim), compile_fn=_OracleBacked)
    assert victim.read_text() == 'keep me'
    svc = service.DecoderService(str(private / 'svc.sock'), compile_fn=_OracleBacked)
    assert stat.S_IMODE(os.stat(str(private / 'svc.sock')).st_mode) == 0o600
    svc._sock.close()
