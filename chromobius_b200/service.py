"""One GPU decoder service shared by many sinter worker processes.

sinter (`sinter.collect(..., num_workers=N)`) starts N worker processes; with the reference each of them compiles its own
chromobius decoder and decodes on its own CPU core (pybind/sinter_compat.pybind.cc:74-147). Here the decoding happens on
a GPU, so N processes would each hold a CUDA context, a copy of the tables and their own scratch. This module lets them
share ONE process that owns the GPU:

    server:   python -m chromobius_b200.service --address /run/user/1000/cb200/sock [--device 0]
    workers:  CHROMOBIUS_B200_SERVICE=/run/user/1000/cb200/sock  ->  chromobius_b200.sinter_decoders()['chromobius']
              talks to the server instead of compiling locally (same sinter.Decoder / sinter.CompiledDecoder surface).

Wire format: raw, length-prefixed binary frames over a Unix stream socket -- no pickle, nothing a peer sends is ever
evaluated. Every frame is `op (u8) | payload length (u64 LE) | payload`:
    request  COMPILE  payload = DEM text (utf-8)          reply OK payload = key (32 B) | num_detectors u64 | num_observables u64
    request  DECODE   payload = key (32 B) | shots u64 | nbytes u64 | shots x nbytes detection-event bytes
                                                          reply OK payload = shots x ceil(num_observables / 8) prediction bytes
    request  STATS    (empty)                             reply OK payload = JSON text
    request  SHUTDOWN (empty)                             reply OK (empty)
    reply    ERROR    payload = message (utf-8)           (the reference's ValueError text for compile / decode failures)
Trust model: whoever can connect can use the GPU and make the server parse a DEM; access control is the file system's.
The socket is created with mode 0600 inside a directory that must not be writable by other users (the server refuses
otherwise and creates a missing directory with mode 0700); it never unlinks a path that is not a socket.

Decoders are cached by the SHA-256 of the DEM text, so N workers decoding the same task compile once. Requests from
different workers are serialised on the one GPU stream (a decoder handle must not be used from two threads at once).

`compile_fn` is injectable so that the CPU test tier can exercise the protocol without a GPU; the product default is the
CUDA decoder and it fails loudly when there is no GPU -- there is no CPU fallback.
"""
import argparse
import hashlib
import json
import os
import socket
import stat
import struct
import threading

import numpy as np

ENV = 'CHROMOBIUS_B200_SERVICE'

OP_COMPILE, OP_DECODE, OP_STATS, OP_SHUTDOWN = 1, 2, 3, 4
REPLY_OK, REPLY_ERROR = 0, 1
MAX_FRAME = 1 << 40


def _recv_exact(sock, n):
    buf = bytearray(n)
    view = memoryview(buf)
    got = 0
    while got < n:
        k = sock.recv_into(view[got:], n - got)
        if k == 0:
            raise EOFError('peer closed the connection')
        got += k
    return buf


def _recv_frame(sock):
    head = _recv_exact(sock, 9)
    op, length = struct.unpack('<BQ', head)
    if length > MAX_FRAME:
        raise ValueError('frame too large')
    return op, (_recv_exact(sock, length) if length else bytearray())


def _send_frame(sock, op, *parts):
    length = sum(len(p) for p in parts)
    sock.sendall(struct.pack('<BQ', op, length))
    for p in parts:
        if len(p):
            sock.sendall(p)


def _prepare_socket_path(address):
    """Refuses unsafe locations; removes a stale socket (and nothing else) at `address`."""
    directory = os.path.dirname(os.path.abspath(address))
    if not os.path.isdir(directory):
        os.makedirs(directory, mode=0o700)
    st = os.stat(directory)
    if st.st_uid != os.getuid() or (st.st_mode & 0o022):
        raise PermissionError('%s must belong to this user and must not be group/world writable' % directory)
    if os.path.lexists(address):
        if not stat.S_ISSOCK(os.lstat(address).st_mode):
            raise FileExistsError('%s exists and is not a socket' % address)
        os.unlink(address)


class DecoderService:
    def __init__(self, address, compile_fn=None, device=None):
        if compile_fn is None:
            import chromobius_b200
            compile_fn = lambda text: chromobius_b200._compile_text(text, device=device)
        self._compile = compile_fn
        _prepare_socket_path(address)
        self._sock = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
        old_umask = os.umask(0o177)
        try:
            self._sock.bind(address)
        finally:
            os.umask(old_umask)
        self._sock.listen(64)
        self._sock.settimeout(0.2)
        self.address = address
        self._decoders = {}
        self._gpu = threading.Lock()       # one decode at a time: one stream, one scratch arena per decoder
        self._stop = threading.Event()
        self._threads = []
        self.stats = {'compiles': 0, 'compile_hits': 0, 'decode_calls': 0, 'shots': 0, 'connections': 0}

    # -- request handlers -------------------------------------------------------------------------------------------
    def _handle(self, op, payload):
        """-> list of byte-like reply parts."""
        if op == OP_COMPILE:
            text = bytes(payload).decode('utf-8')
            key = hashlib.sha256(bytes(payload)).digest()
            with self._gpu:
                dec = self._decoders.get(key)
                if dec is None:
                    dec = self._compile(text)
                    self._decoders[key] = dec
                    self.stats['compiles'] += 1
                else:
                    self.stats['compile_hits'] += 1
            return [key, struct.pack('<QQ', int(dec.num_detectors), int(dec.num_observables))]
        if op == OP_DECODE:
            if len(payload) < 48:
                raise ValueError('short decode request')
            key = bytes(payload[:32])
            shots, nbytes = struct.unpack('<QQ', payload[32:48])
            dec = self._decoders.get(key)
            if dec is None:
                raise ValueError('unknown decoder key (compile first)')
            if len(payload) - 48 != shots * nbytes:
                raise ValueError('decode request holds %d bytes, expected %d x %d' % (len(payload) - 48, shots, nbytes))
            dets = np.frombuffer(payload, dtype=np.uint8, offset=48).reshape(shots, nbytes)
            with self._gpu:
                pred = dec.predict_obs_flips_from_dets_bit_packed(dets)
                self.stats['decode_calls'] += 1
                self.stats['shots'] += int(shots)
            return [np.ascontiguousarray(pred, dtype=np.uint8).tobytes()]
        if op == OP_STATS:
            return [json.dumps(dict(self.stats, decoders=len(self._decoders))).encode()]
        if op == OP_SHUTDOWN:
            self._stop.set()
            return []
        raise ValueError('unknown request %r' % (op,))

    def _serve_connection(self, conn):
        try:
            while not self._stop.is_set():
                try:
                    op, payload = _recv_frame(conn)
                except (EOFError, OSError, ValueError):
                    break
                try:
                    parts = self._handle(op, payload)
                    _send_frame(conn, REPLY_OK, *parts)
                except Exception as ex:  # compile / decode errors travel back as the reference's ValueError text
                    _send_frame(conn, REPLY_ERROR, ('%s' % (ex,)).encode('utf-8'))
                if op == OP_SHUTDOWN:
                    break
        finally:
            conn.close()

    def serve_forever(self):
        """Accepts workers until a SHUTDOWN request arrives."""
        while not self._stop.is_set():
            try:
                conn, _ = self._sock.accept()
            except socket.timeout:  # look at the stop flag again
                continue
            except OSError:
                continue
            conn.settimeout(None)
            self.stats['connections'] += 1
            t = threading.Thread(target=self._serve_connection, args=(conn,), daemon=True)
            t.start()
            self._threads.append(t)
        self._sock.close()
        try:
            os.unlink(self.address)
        except OSError:
            pass
        for t in self._threads:
            t.join(timeout=1.0)

    def start_in_thread(self):
        t = threading.Thread(target=self.serve_forever, daemon=True)
        t.start()
        return t


class RemoteCompiledDecoder:
    """Client side: the sinter.CompiledDecoder surface (pybind/sinter_compat.pybind.cc:23-72) over the service."""

    def __init__(self, address, dem_text):
        self._conn = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
        self._conn.connect(address)
        reply = self._call(OP_COMPILE, dem_text.encode('utf-8'))
        self._key = bytes(reply[:32])
        self.num_detectors, self.num_observables = struct.unpack('<QQ', reply[32:48])
        self.num_detector_bytes = (self.num_detectors + 7) // 8
        self.num_observable_bytes = (self.num_observables + 7) // 8

    def _call(self, op, *parts):
        _send_frame(self._conn, op, *parts)
        status, payload = _recv_frame(self._conn)
        if status != REPLY_OK:
            raise ValueError(bytes(payload).decode('utf-8', 'replace'))
        return payload

    def predict_obs_flips_from_dets_bit_packed(self, dets):
        dets = np.ascontiguousarray(dets, dtype=np.uint8)
        if dets.ndim != 2 or dets.shape[1] != self.num_detector_bytes:
            raise ValueError('dets.shape[1] != num_detector_bytes')
        reply = self._call(OP_DECODE, self._key, struct.pack('<QQ', int(dets.shape[0]), int(dets.shape[1])), memoryview(dets).cast('B'))
        return np.frombuffer(reply, dtype=np.uint8).reshape(dets.shape[0], self.num_observable_bytes).copy()

    def decode_shots_bit_packed(self, *, bit_packed_detection_event_data):
        d = bit_packed_detection_event_data
        if d.ndim != 2:
            raise ValueError('bit_packed_detection_event_data.ndim() != 2')
        if d.shape[1] != self.num_detector_bytes:
            raise ValueError('bit_packed_detection_event_data.shape(1) != num_detector_bytes')
        return self.predict_obs_flips_from_dets_bit_packed(d)

    def service_stats(self):
        return json.loads(bytes(self._call(OP_STATS)).decode())

    def shutdown_service(self):
        self._call(OP_SHUTDOWN)

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
    ap.add_argument('--address', required=True, help='path of the Unix socket to listen on (in a directory only you can write)')
    ap.add_argument('--device', type=int, default=None)
    args = ap.parse_args(argv)
    svc = DecoderService(args.address, device=args.device)
    print('chromobius_b200 service listening on %s' % args.address, flush=True)
    svc.serve_forever()


if __name__ == '__main__':
    main()
