"""One GPU decoder service shared by many sinter worker processes.

sinter (`sinter.collect(..., num_workers=N)`) starts N worker processes; with the reference each of them compiles its own
chromobius decoder and decodes on its own CPU core (pybind/sinter_compat.pybind.cc:74-147). Here the decoding happens on
a GPU, so N processes would each hold a CUDA context, a copy of the tables and their own scratch. This module lets them
share ONE process that owns the GPU:

    server:   python -m chromobius_b200.service --address /tmp/cb200.sock [--device 0]
    workers:  CHROMOBIUS_B200_SERVICE=/tmp/cb200.sock  ->  chromobius_b200.sinter_decoders()['chromobius'] talks to the
              server instead of compiling locally (same sinter.Decoder / sinter.CompiledDecoder surface).

Protocol (multiprocessing.connection over a Unix socket, one connection per worker, requests served in arrival order):
    ('compile', dem_text)                 -> ('ok', key, num_detectors, num_observables) | ('error', message)
    ('decode', key, shots, nbytes, raw)   -> ('ok', raw predictions)                     | ('error', message)
    ('stats',)                            -> ('ok', {...})
    ('shutdown',)                         -> ('ok',)
Decoders are cached by a hash of the DEM text, so N workers decoding the same task compile once. Requests from different
workers are serialised on the one GPU stream (a decoder handle must not be used from two threads at once).

`compile_fn` is injectable so that the CPU test tier can exercise the protocol without a GPU; the product default is the
CUDA decoder and it fails loudly when there is no GPU -- there is no CPU fallback.
"""
import argparse
import hashlib
import os
import threading
from multiprocessing.connection import Client, Listener

import numpy as np

ENV = 'CHROMOBIUS_B200_SERVICE'


class DecoderService:
    def __init__(self, address, compile_fn=None, device=None, authkey=b'chromobius_b200'):
        if compile_fn is None:
            import chromobius_b200
            compile_fn = lambda text: chromobius_b200._compile_text(text, device=device)
        self._compile = compile_fn
        self._listener = Listener(address, family='AF_UNIX', authkey=authkey)
        self.address = address
        self._decoders = {}
        self._gpu = threading.Lock()       # one decode at a time: one stream, one scratch arena per decoder
        self._stop = threading.Event()
        self._threads = []
        self.stats = {'compiles': 0, 'compile_hits': 0, 'decode_calls': 0, 'shots': 0, 'connections': 0}

    # -- request handlers -------------------------------------------------------------------------------------------
    def _handle(self, msg):
        op = msg[0]
        if op == 'compile':
            text = msg[1]
            key = hashlib.sha256(text.encode('utf-8')).hexdigest()
            with self._gpu:
                dec = self._decoders.get(key)
                if dec is None:
                    dec = self._compile(text)
                    self._decoders[key] = dec
                    self.stats['compiles'] += 1
                else:
                    self.stats['compile_hits'] += 1
            return ('ok', key, int(dec.num_detectors), int(dec.num_observables))
        if op == 'decode':
            _, key, shots, nbytes, raw = msg
            dec = self._decoders.get(key)
            if dec is None:
                return ('error', 'unknown decoder key (compile first)')
            dets = np.frombuffer(raw, dtype=np.uint8)
            if dets.size != shots * nbytes:
                return ('error', 'decode request holds %d bytes, expected %d x %d' % (dets.size, shots, nbytes))
            dets = dets.reshape(shots, nbytes)
            with self._gpu:
                pred = dec.predict_obs_flips_from_dets_bit_packed(dets)
                self.stats['decode_calls'] += 1
                self.stats['shots'] += int(shots)
            return ('ok', np.ascontiguousarray(pred, dtype=np.uint8).tobytes())
        if op == 'stats':
            return ('ok', dict(self.stats, decoders=len(self._decoders)))
        if op == 'shutdown':
            self._stop.set()
            return ('ok',)
        return ('error', 'unknown request %r' % (op,))

    def _serve_connection(self, conn):
        try:
            while not self._stop.is_set():
                try:
                    msg = conn.recv()
                except (EOFError, OSError):
                    break
                try:
                    reply = self._handle(msg)
                except Exception as ex:  # compile / decode errors travel back as the reference's ValueError text
                    reply = ('error', '%s' % (ex,))
                conn.send(reply)
                if msg and msg[0] == 'shutdown':
                    break
        finally:
            conn.close()

    def serve_forever(self):
        """Accepts workers until a 'shutdown' request arrives."""
        self._listener._listener._socket.settimeout(0.2)
        while not self._stop.is_set():
            try:
                conn = self._listener.accept()
            except OSError:  # timeout: look at the stop flag again
                continue
            except Exception:
                continue
            self.stats['connections'] += 1
            t = threading.Thread(target=self._serve_connection, args=(conn,), daemon=True)
            t.start()
            self._threads.append(t)
        self._listener.close()
        for t in self._threads:
            t.join(timeout=1.0)

    def start_in_thread(self):
        t = threading.Thread(target=self.serve_forever, daemon=True)
        t.start()
        return t


class RemoteCompiledDecoder:
    """Client side: the sinter.CompiledDecoder surface (pybind/sinter_compat.pybind.cc:23-72) over the service."""

    def __init__(self, address, dem_text, authkey=b'chromobius_b200'):
        self._conn = Client(address, family='AF_UNIX', authkey=authkey)
        reply = self._call(('compile', dem_text))
        self._key, self.num_detectors, self.num_observables = reply[1], reply[2], reply[3]
        self.num_detector_bytes = (self.num_detectors + 7) // 8
        self.num_observable_bytes = (self.num_observables + 7) // 8

    def _call(self, msg):
        self._conn.send(msg)
        reply = self._conn.recv()
        if reply[0] != 'ok':
            raise ValueError(reply[1])
        return reply

    def predict_obs_flips_from_dets_bit_packed(self, dets):
        dets = np.ascontiguousarray(dets, dtype=np.uint8)
        if dets.ndim != 2 or dets.shape[1] != self.num_detector_bytes:
            raise ValueError('dets.shape[1] != num_detector_bytes')
        reply = self._call(('decode', self._key, int(dets.shape[0]), int(dets.shape[1]), dets.tobytes()))
        return np.frombuffer(reply[1], dtype=np.uint8).reshape(dets.shape[0], self.num_observable_bytes).copy()

    def decode_shots_bit_packed(self, *, bit_packed_detection_event_data):
        d = bit_packed_detection_event_data
        if d.ndim != 2:
            raise ValueError('bit_packed_detection_event_data.ndim() != 2')
        if d.shape[1] != self.num_detector_bytes:
            raise ValueError('bit_packed_detection_event_data.shape(1) != num_detector_bytes')
        return self.predict_obs_flips_from_dets_bit_packed(d)

    def service_stats(self):
        return self._call(('stats',))[1]

    def shutdown_service(self):
        self._call(('shutdown',))

    def close(self):
        try:
            self._conn.close()
        except Exception:
            pass


def service_address():
    """Address of the shared decoder service this process should use, or None (environment CHROMOBIUS_B200_SERVICE)."""
    a = os.environ.get(ENV)
    return a if a else None


def main(argv=None):
    ap = argparse.ArgumentParser(description='GPU decoder service shared by sinter worker processes')
    ap.add_argument('--address', required=True, help='path of the Unix socket to listen on')
    ap.add_argument('--device', type=int, default=None)
    args = ap.parse_args(argv)
    if os.path.exists(args.address):
        os.unlink(args.address)
    svc = DecoderService(args.address, device=args.device)
    print('chromobius_b200 service listening on %s' % args.address, flush=True)
    svc.serve_forever()


if __name__ == '__main__':
    main()
