"""A from-scratch, dependency-free implementation of the slice of the ``pynng`` API the
DetectMate service uses, speaking the real NNG SP/PAIR0 wire protocol.

Why: the reference's data channel is ``pynng.Pair0`` (libnng through cffi;
/root/reference/src/service/features/engine.py:2,121-131,163,211,234 and
features/engine_socket.py:32-55); neither pynng nor libnng can be installed offline.  This
module is put on ``sys.path`` by ``detectmateservice_b200.compat.install_shims()`` only when
the real package is missing.  Because it speaks the standard wire format it interoperates
with real NNG peers (fluentd's nng plugins, the reference's own services):

  ipc:///path   AF_UNIX stream socket;  tcp://host:port  TCP;  inproc://name  in-process
  handshake     both sides send  00 'S' 'P' 00 | proto (be16) | 00 00 ; PAIR v0 = 0x0010
  frames        tcp:  len (be64) | payload        ipc:  0x01 | len (be64) | payload

PAIR0 semantics reproduced: one peer at a time; ``dial(block=False)`` keeps retrying in
the background and reconnects; ``send(block=False)`` hands the message to the transport at
once when the kernel takes it, otherwise to a bounded per-connection send queue drained by a
writer thread, and raises ``TryAgain`` only without a connected peer or when that queue is
full; ``recv`` honours ``recv_timeout`` (ms) with ``Timeout``; a frame larger than
``recv_max_size`` closes the connection (NNG's NNG_OPT_RECVMAXSZ; 0 = unlimited, which is the
default here because the reference's own tests push 1 MiB + 11 bytes through default sockets,
/root/reference/tests/test_engine_multi_output.py:429-447); using a closed socket raises
``Closed`` (an ``NNGException``).
"""
from __future__ import annotations

import collections
import errno
import os
import queue
import select
import socket
import struct
import threading
import time
from typing import Dict, Optional
from urllib.parse import urlparse

__shim__ = True
__version__ = "0.9.0+b200shim"


class NNGException(Exception):
    pass


class Timeout(NNGException):
    pass


class TryAgain(NNGException):
    pass


class Closed(NNGException):
    pass


class AddressInUse(NNGException):
    pass


class ConnectionRefused(NNGException):
    pass


class NotSupported(NNGException):
    pass


class BadScheme(NNGException):
    pass


class exceptions:  # namespace mirror: pynng.exceptions.AddressInUse ...
    NNGException = NNGException
    Timeout = Timeout
    TryAgain = TryAgain
    Closed = Closed
    AddressInUse = AddressInUse
    ConnectionRefused = ConnectionRefused
    NotSupported = NotSupported
    BadScheme = BadScheme


_PAIR0 = 0x0010
_HANDSHAKE = b"\x00SP\x00" + struct.pack(">H", _PAIR0) + b"\x00\x00"
_RECONNECT_S = 0.05
_SENDQ_MSGS = 256                 # messages a connection may hold for its writer thread ...
_SENDQ_BYTES = 256 << 20          # ... and their bytes (whichever is reached first)

_inproc_lock = threading.Lock()
_inproc_listeners: Dict[str, "Socket"] = {}


def _recv_exact(sock: socket.socket, n: int) -> Optional[bytes]:
    chunks = []
    while n:
        try:
            b = sock.recv(min(n, 1 << 20))
        except OSError:
            return None
        if not b:
            return None
        chunks.append(b)
        n -= len(b)
    return b"".join(chunks)


def _recv_into(sock: socket.socket, mv: memoryview) -> Optional[memoryview]:
    n, got = len(mv), 0
    while got < n:
        try:
            k = sock.recv_into(mv[got:], n - got)
        except OSError:
            return None
        if k == 0:
            return None
        got += k
    return mv


def _recv_into_new(sock: socket.socket, n: int) -> Optional[bytearray]:
    """Large frames: one allocation, filled in place (no per-chunk objects, no join copy)."""
    buf = bytearray(n)
    mv = memoryview(buf)
    got = 0
    while got < n:
        try:
            k = sock.recv_into(mv[got:], n - got)
        except OSError:
            return None
        if k == 0:
            return None
        got += k
    return buf


class _Pipe:
    """One established connection (after the SP handshake)."""

    def __init__(self, owner: "Socket", sock: Optional[socket.socket], ipc: bool, peer: Optional["_Pipe"] = None):
        self.owner, self.sock, self.ipc = owner, sock, ipc
        self.peer = peer                 # inproc: the other end
        self.alive = True
        self.wlock = threading.Condition()          # guards the socket's write side and the send queue
        self.sendq = collections.deque()            # memoryviews waiting for the writer thread
        self.sendq_bytes = 0
        self.writing = False                        # the writer thread is inside sendall()
        self.writer = None
        if sock is not None:
            threading.Thread(target=self._reader, name="sp-pipe-reader", daemon=True).start()

    def _reader(self) -> None:
        s = self.sock
        while self.alive:
            if self.ipc:
                t = _recv_exact(s, 1)
                if t is None or t != b"\x01":
                    break
            h = _recv_exact(s, 8)
            if h is None:
                break
            (ln,) = struct.unpack(">Q", h)
            rmax = int(getattr(self.owner, "recv_max_size", 0) or 0)
            if rmax and ln > rmax:
                break                              # NNG: an oversized message closes the pipe
            alloc = getattr(self.owner, "frame_allocator", None)
            frame = alloc(ln) if (alloc is not None and ln >= 65536) else None
            if frame is not None:
                # opt-in (shim only): the consumer lends receive buffers (e.g. pinned host memory
                # the GPU can DMA from); the frame is delivered as a memoryview into that buffer
                body = _recv_into(s, memoryview(frame)[:ln])
            elif ln >= 65536 and getattr(self.owner, "large_frames_as_bytearray", False):
                body = _recv_into_new(s, ln)       # opt-in (shim only): bytes-like, not bytes
            else:
                body = _recv_exact(s, ln) if ln else b""
            if body is None:
                break
            self.owner._deliver(body)
        self.close()

    def _frame(self, data) -> list:
        hdr = (b"\x01" if self.ipc else b"") + struct.pack(">Q", len(data))
        return [memoryview(hdr + bytes(data))] if len(data) < 65536 else [memoryview(hdr), memoryview(data)]

    def _writer(self) -> None:
        while True:
            with self.wlock:
                while self.alive and not self.sendq:
                    self.writing = False
                    self.wlock.notify_all()
                    self.wlock.wait(0.2)
                if not self.alive:
                    self.writing = False
                    self.wlock.notify_all()
                    return
                mv = self.sendq.popleft()
                self.sendq_bytes -= len(mv)
                self.writing = True
            try:
                self.sock.sendall(mv)
            except OSError:
                self.close()
                return

    def send_nowait(self, data) -> bool:
        """Hand one message to the transport without waiting.  False = nothing can take it now."""
        if self.sock is None:                      # inproc
            if self.peer is None or not self.peer.alive:
                raise Closed("peer gone")
            self.peer.owner._deliver(bytes(data))
            return True
        parts = self._frame(data)
        with self.wlock:
            if not self.alive:
                return False
            if not self.sendq and not self.writing:
                # nothing queued: try the kernel directly (it takes what fits into the socket buffer)
                try:
                    while parts:
                        n = self.sock.send(parts[0], socket.MSG_DONTWAIT)
                        if n < len(parts[0]):
                            parts[0] = parts[0][n:]
                            break
                        parts.pop(0)
                except (BlockingIOError, InterruptedError):
                    pass
                except OSError as e:
                    self.alive and self.close_locked_error(e)
                    raise Closed(str(e)) from e
                if not parts:
                    return True
                started = True
            else:
                started = False
                if len(self.sendq) >= _SENDQ_MSGS or self.sendq_bytes + sum(len(p) for p in parts) > _SENDQ_BYTES:
                    return False
            # (a frame that has started must be finished: it goes to the queue even when the queue is long)
            _ = started
            for p_ in parts:
                self.sendq.append(p_)
                self.sendq_bytes += len(p_)
            if self.writer is None:
                self.writer = threading.Thread(target=self._writer, name="sp-pipe-writer", daemon=True)
                self.writer.start()
            self.wlock.notify_all()
            return True

    def close_locked_error(self, e) -> None:
        threading.Thread(target=self.close, daemon=True).start()

    def send(self, data, deadline: Optional[float] = None) -> None:
        """Blocking send: behind whatever is queued, then straight to the socket."""
        if self.sock is None:                      # inproc
            if self.peer is None or not self.peer.alive:
                raise Closed("peer gone")
            self.peer.owner._deliver(bytes(data))
            return
        parts = self._frame(data)
        with self.wlock:
            while self.alive and (self.sendq or self.writing):
                if deadline is not None and time.monotonic() >= deadline:
                    raise Timeout("send timed out")
                self.wlock.wait(0.05)
            if not self.alive:
                raise Closed("connection closed")
            try:
                for p_ in parts:
                    self.sock.sendall(p_)
            except OSError as e:
                self.close_locked_error(e)
                raise Closed(str(e)) from e

    def close(self) -> None:
        if not self.alive:
            return
        self.alive = False
        try:
            with self.wlock:
                self.wlock.notify_all()
        except RuntimeError:
            pass
        if self.sock is not None:
            try:
                self.sock.shutdown(socket.SHUT_RDWR)
            except OSError:
                pass
            try:
                self.sock.close()
            except OSError:
                pass
        if self.peer is not None and self.peer.alive:
            p, self.peer = self.peer, None
            p.peer = None
            p.close()
        self.owner._pipe_closed(self)


class Socket:
    """Base of the SP sockets (only PAIR0 is implemented)."""

    def __init__(self, listen: Optional[str] = None, dial: Optional[str] = None, recv_timeout: Optional[int] = None,
                 send_timeout: Optional[int] = None, recv_buffer_size: int = 128, send_buffer_size: int = 128,
                 block_on_dial: Optional[bool] = None, recv_max_size: int = 0, **_ignored) -> None:
        self.recv_timeout = recv_timeout           # ms, None / negative = wait forever
        self.send_timeout = send_timeout
        self.dial_timeout = 1000
        self.recv_buffer_size = recv_buffer_size
        self.send_buffer_size = send_buffer_size
        self.recv_max_size = recv_max_size         # bytes; larger frames close the connection (0 = unlimited)
        self._rx: "queue.Queue[bytes]" = queue.Queue()
        self._pipe: Optional[_Pipe] = None
        self._pipe_cv = threading.Condition()
        self._closed = False
        self._listen_socks = []
        self._unlink_paths = []
        self._threads = []
        self._inproc_names = []
        if listen:
            self.listen(listen)
        if dial:
            self.dial(dial, block=block_on_dial)

    # ------------------------------------------------------------------ context manager
    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False

    # ------------------------------------------------------------------ internals
    def _deliver(self, body: bytes) -> None:
        # NNG holds at most recv_buffer_size messages per socket; beyond that the pipe's reader
        # stalls and the transport (TCP / unix socket buffers) pushes back on the sender.
        cap = max(1, int(self.recv_buffer_size or 0))
        while self._rx.qsize() >= cap and not self._closed:
            time.sleep(0.0002)
        self._rx.put(body)

    def _adopt(self, pipe: _Pipe) -> bool:
        with self._pipe_cv:
            if self._closed or (self._pipe is not None and self._pipe.alive):
                return False
            self._pipe = pipe
            self._pipe_cv.notify_all()
            return True

    def _pipe_closed(self, pipe: _Pipe) -> None:
        with self._pipe_cv:
            if self._pipe is pipe:
                self._pipe = None
            self._pipe_cv.notify_all()

    @staticmethod
    def _handshake(s: socket.socket) -> bool:
        try:
            s.settimeout(2.0)
            s.sendall(_HANDSHAKE)
            got = _recv_exact(s, 8)
            s.settimeout(None)
        except OSError:
            return False
        return got is not None and got[:4] == b"\x00SP\x00" and struct.unpack(">H", got[4:6])[0] == _PAIR0

    @staticmethod
    def _parse(addr: str):
        u = urlparse(addr)
        if u.scheme == "ipc":
            path = addr[len("ipc://"):]
            if not path:
                raise BadScheme(f"empty ipc path in {addr!r}")
            return "ipc", path
        if u.scheme == "tcp":
            if not u.port:
                raise BadScheme(f"missing port in {addr!r}")
            host = u.hostname or "127.0.0.1"
            return "tcp", (("0.0.0.0" if host == "*" else host), int(u.port))
        if u.scheme == "inproc":
            return "inproc", addr
        # NNG answers NNG_ENOTSUP for a transport it does not know (pynng.exceptions.NotSupported,
        # /root/reference/tests/test_engine_socket_factory_error_handling.py:96-101)
        raise NotSupported(f"transport {u.scheme!r} is not supported ({addr!r})")

    # ------------------------------------------------------------------ listen / dial
    def listen(self, addr: str, flags: int = 0) -> None:
        if self._closed:
            raise Closed("socket is closed")
        kind, where = self._parse(addr)
        if kind == "inproc":
            with _inproc_lock:
                if where in _inproc_listeners:
                    raise AddressInUse(addr)
                _inproc_listeners[where] = self
            self._inproc_names.append(where)
            return
        ls = socket.socket(socket.AF_UNIX if kind == "ipc" else socket.AF_INET, socket.SOCK_STREAM)
        try:
            if kind == "tcp":
                ls.setsockopt(socket.SOL_SOCKET, socket.SO_REUSEADDR, 1)
            ls.bind(where)
            ls.listen(8)
        except OSError as e:
            ls.close()
            if e.errno in (errno.EADDRINUSE,):
                raise AddressInUse(f"{addr}: {e}") from e
            raise NNGException(f"cannot listen on {addr}: {e}") from e
        if kind == "ipc":
            self._unlink_paths.append(where)
        self._listen_socks.append(ls)
        t = threading.Thread(target=self._acceptor, args=(ls, kind == "ipc"), name="sp-acceptor", daemon=True)
        t.start()
        self._threads.append(t)

    def _acceptor(self, ls: socket.socket, ipc: bool) -> None:
        while not self._closed:
            try:
                c, _ = ls.accept()
            except OSError:
                return
            if not ipc:
                try:
                    c.setsockopt(socket.IPPROTO_TCP, socket.TCP_NODELAY, 1)
                except OSError:
                    pass
            if not self._handshake(c):
                c.close()
                continue
            with self._pipe_cv:
                busy = self._pipe is not None and self._pipe.alive
            if busy or self._closed:               # PAIR0: one peer at a time
                c.close()
                continue
            self._adopt(_Pipe(self, c, ipc))

    def dial(self, addr: str, block: Optional[bool] = None) -> None:
        if self._closed:
            raise Closed("socket is closed")
        kind, where = self._parse(addr)
        if block is None or block:
            ok = self._connect_once(kind, where)
            if ok:
                self._start_redialer(kind, where)
                return
            if block:
                raise ConnectionRefused(addr)
        self._start_redialer(kind, where)

    def _connect_once(self, kind: str, where) -> bool:
        if kind == "inproc":
            with _inproc_lock:
                peer = _inproc_listeners.get(where)
            if peer is None or peer._closed:
                return False
            a = _Pipe(self, None, False)
            b = _Pipe(peer, None, False, peer=a)
            a.peer = b
            if not peer._adopt(b):
                a.alive = b.alive = False
                return False
            if not self._adopt(a):
                b.close()
                return False
            return True
        s = socket.socket(socket.AF_UNIX if kind == "ipc" else socket.AF_INET, socket.SOCK_STREAM)
        try:
            s.settimeout(max(self.dial_timeout, 1) / 1000.0)
            s.connect(where)
            s.settimeout(None)
            if kind == "tcp":
                s.setsockopt(socket.IPPROTO_TCP, socket.TCP_NODELAY, 1)
        except OSError:
            s.close()
            return False
        if not self._handshake(s):
            s.close()
            return False
        if not self._adopt(_Pipe(self, s, kind == "ipc")):
            s.close()
            return False
        return True

    def _start_redialer(self, kind: str, where) -> None:
        def loop():
            while not self._closed:
                with self._pipe_cv:
                    connected = self._pipe is not None and self._pipe.alive
                if not connected:
                    self._connect_once(kind, where)
                time.sleep(_RECONNECT_S)
        t = threading.Thread(target=loop, name="sp-dialer", daemon=True)
        t.start()
        self._threads.append(t)

    # ------------------------------------------------------------------ data
    def recv(self, block: bool = True) -> bytes:
        if self._closed:
            raise Closed("socket is closed")
        if not block:
            try:
                return self._rx.get_nowait()
            except queue.Empty:
                raise TryAgain("no message") from None
        tmo = self.recv_timeout
        deadline = None if tmo is None or tmo < 0 else time.monotonic() + tmo / 1000.0
        while True:
            if self._closed:
                raise Closed("socket is closed")
            wait = 0.05 if deadline is None else min(0.05, max(0.0, deadline - time.monotonic()))
            try:
                return self._rx.get(timeout=wait) if wait > 0 else self._rx.get_nowait()
            except queue.Empty:
                if deadline is not None and time.monotonic() >= deadline:
                    raise Timeout("recv timed out") from None

    def send(self, data: bytes, block: bool = True) -> None:
        if self._closed:
            raise Closed("socket is closed")
        if not isinstance(data, (bytes, bytearray, memoryview)):
            data = bytes(data)
        if not block:
            with self._pipe_cv:
                p = self._pipe
            if p is None or not p.alive or not p.send_nowait(data):
                raise TryAgain("no peer ready")
            return
        tmo = self.send_timeout
        deadline = None if tmo is None or tmo < 0 else time.monotonic() + tmo / 1000.0
        with self._pipe_cv:
            while not self._closed and (self._pipe is None or not self._pipe.alive):
                remaining = None if deadline is None else deadline - time.monotonic()
                if remaining is not None and remaining <= 0:
                    raise Timeout("send timed out")
                self._pipe_cv.wait(0.05 if remaining is None else min(0.05, remaining))
            if self._closed:
                raise Closed("socket is closed")
            p = self._pipe
        p.send(data, deadline)

    # ------------------------------------------------------------------ teardown
    def close(self) -> None:
        if self._closed:
            return
        self._closed = True
        for name in self._inproc_names:
            with _inproc_lock:
                if _inproc_listeners.get(name) is self:
                    del _inproc_listeners[name]
        for ls in self._listen_socks:
            try:
                ls.close()
            except OSError:
                pass
        with self._pipe_cv:
            p = self._pipe
            self._pipe_cv.notify_all()
        if p is not None:
            p.close()
        for path in self._unlink_paths:
            try:
                os.unlink(path)
            except OSError:
                pass


class Pair0(Socket):
    pass
