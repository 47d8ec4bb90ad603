"""A small explicit-state model of the peer-memory halo exchange protocol (csrc/p2p.cu, api.cu:
levelAdvance / pushHalo / haloWait), explored over ALL interleavings of the ranks' streams.

The device code cannot run here; what can be checked on a CPU is the ARGUMENT the code relies on
(p2p.cu header): with symmetric peer sets
  * no "ready to receive" handshake is needed for the double-buffered state -- a peer can store
    exchange n+1 into the buffer a rank reads during sweep n only after that sweep has finished;
  * the single-buffered auxField rows DO need one (`ready[]`);
  * none of the three placements of the wait (wait kernel behind the push, wait inside the next
    sweep by its halo CTAs, push on a second stream with the halo CTAs appended to the sweep)
    can deadlock or read stale / half-written halo rows, and in the two-stream variant the sweep
    of step n+2 cannot overwrite the buffer the push of step n is still reading;
and that the checker has teeth: with a one-directional peer relation -- which musb200_p2p_connect
rejects (MUSB200_ERR_UNSUPPORTED, the rank stays on NCCL) -- or without the auxField handshake,
the same exploration FINDS the race.

Model: per rank one or two in-order streams of atomic events; a sweep / push is a begin and an end
event, reads and writes last from one to the other.  Step n reads buffer (n+1) % 2 and writes
buffer n % 2 (mus_swap_Now_Next); push n stores this rank's links of buffer n % 2 into the halo
rows of the same buffer on every peer, then publishes n into arrived[me] there."""
import pytest


class Violation(Exception):
    pass


def build_programs(R, K, send, mode, aux, handshake, no_evpushed=False):
    """send[r] = ranks r stores to; recv[r] = ranks that store to r.  Returns streams: list of
    (rank, [events])"""
    streams = []
    for r in range(R):
        main, comm = [], []
        for n in range(1, K + 1):
            if mode == "wait-kernel":
                main += [("sweep_begin", n), ("halo_begin", n), ("halo_end", n), ("sweep_end", n)]
                if aux and handshake:
                    main += [("ready_pub", n), ("ready_wait", n)]
                main += [("push_begin", n), ("push_end", n), ("publish", n), ("wait", n)]
                if aux:
                    main += [("aux_read_begin", n), ("aux_read_end", n)]       # interpolation reads aux halos
            elif mode == "sweep-wait":
                main += [("sweep_begin", n)]
                if n > 1:
                    main += [("wait", n - 1)]
                main += [("halo_begin", n), ("halo_end", n), ("sweep_end", n),
                         ("push_begin", n), ("push_end", n), ("publish", n)]
            elif mode == "overlap":
                main += ([] if no_evpushed else [("wait_pushed", n)]) + [("sweep_begin", n)]
                if n > 1:
                    main += [("wait", n - 1)]
                main += [("halo_begin", n), ("halo_end", n), ("sweep_end", n), ("rec_swept", n)]
                comm += [("wait_swept", n), ("push_begin", n), ("push_end", n), ("publish", n), ("rec_pushed", n)]
            else:
                raise ValueError(mode)
        if mode != "wait-kernel":
            main += [("wait", K)]                                              # musb200_synchronize
        streams.append((r, main))
        if comm:
            streams.append((r, comm))
    return streams


def explore(R, K, send, mode, aux=False, handshake=True, no_evpushed=False):
    """exhaustive DFS over the interleavings; raises Violation on the first property broken,
    returns the number of distinct states otherwise"""
    recv = [[p for p in range(R) if r in send[p]] for r in range(R)]
    streams = build_programs(R, K, send, mode, aux, handshake, no_evpushed)
    S = len(streams)
    # shared state, all small ints in flat tuples:
    #  arrived[r][p], ready[r][p], tag[r][b][p], writing[r][b][p], reading[r][b], auxtag[r][p],
    #  auxwriting[r][p], auxreading[r], swept[r], pushed[r][b], pushreading[r][b], sweepwriting[r][b]
    def init():
        return dict(pc=[0] * S, arrived=[[0] * R for _ in range(R)], ready=[[0] * R for _ in range(R)],
                    tag=[[[0] * R, [0] * R] for _ in range(R)], writing=[[[0] * R, [0] * R] for _ in range(R)],
                    reading=[[0, 0] for _ in range(R)], auxtag=[[0] * R for _ in range(R)],
                    auxwriting=[[0] * R for _ in range(R)], auxreading=[0] * R, swept=[0] * R,
                    pushed=[[0, 0] for _ in range(R)], pushreading=[[0, 0] for _ in range(R)],
                    sweepwriting=[[0, 0] for _ in range(R)])

    def freeze(x):
        return tuple(freeze(v) for v in x) if isinstance(x, list) else x

    def key(st):
        return tuple(freeze(st[k]) for k in sorted(st))

    def clone(st):
        def c(x):
            return [c(v) for v in x] if isinstance(x, list) else x
        return {k: c(v) for k, v in st.items()}

    def enabled(st, r, ev):
        kind, n = ev
        if kind == "wait":
            return all(st["arrived"][r][p] >= n for p in recv[r])
        if kind == "ready_wait":
            return all(st["ready"][r][p] >= n for p in send[r])
        if kind == "wait_swept":
            return st["swept"][r] >= n
        if kind == "wait_pushed":                       # the push that last read buffer n % 2: step n - 2
            return n <= 2 or st["pushed"][r][n % 2] >= n - 2
        return True

    def apply(st, r, ev):
        kind, n = ev
        rd, wr = (n + 1) % 2, n % 2
        if kind == "sweep_begin":
            if st["pushreading"][r][wr]:
                raise Violation("rank %d: sweep %d overwrites the buffer a push still reads" % (r, n))
            st["sweepwriting"][r][wr] = 1
        elif kind == "sweep_end":
            st["sweepwriting"][r][wr] = 0
        elif kind in ("halo_begin", "halo_end"):
            for p in recv[r]:
                if st["writing"][r][rd][p]:
                    raise Violation("rank %d sweep %d reads halo rows rank %d is writing" % (r, n, p))
                if st["tag"][r][rd][p] != n - 1:
                    raise Violation("rank %d sweep %d reads exchange %d of rank %d, expected %d"
                                    % (r, n, st["tag"][r][rd][p], p, n - 1))
            st["reading"][r][rd] = 1 if kind == "halo_begin" else 0
        elif kind == "push_begin":
            if st["sweepwriting"][r][wr]:
                raise Violation("rank %d: push %d reads a buffer a sweep is writing" % (r, n))
            st["pushreading"][r][wr] = 1
            for p in send[r]:
                if st["reading"][p][wr]:
                    raise Violation("rank %d push %d stores into halo rows rank %d is reading" % (r, n, p))
                st["writing"][p][wr][r] = 1
                if aux:
                    if st["auxreading"][p]:
                        raise Violation("rank %d push %d stores auxField rows rank %d is reading" % (r, n, p))
                    st["auxwriting"][p][r] = 1
        elif kind == "push_end":
            st["pushreading"][r][wr] = 0
            for p in send[r]:
                if st["reading"][p][wr]:
                    raise Violation("rank %d push %d stored into halo rows rank %d is reading" % (r, n, p))
                st["writing"][p][wr][r] = 0
                st["tag"][p][wr][r] = n
                if aux:
                    if st["auxreading"][p]:
                        raise Violation("rank %d push %d stored auxField rows rank %d is reading" % (r, n, p))
                    st["auxwriting"][p][r] = 0
                    st["auxtag"][p][r] = n
        elif kind == "publish":
            for p in send[r]:
                st["arrived"][p][r] = n
        elif kind == "ready_pub":
            for p in send[r]:                            # symmetric sets: my receivers are my senders
                st["ready"][p][r] = n
        elif kind in ("aux_read_begin", "aux_read_end"):
            for p in recv[r]:
                if st["auxwriting"][r][p] or st["auxtag"][r][p] != n:
                    raise Violation("rank %d reads auxField halo rows of exchange %d while rank %d has %s"
                                    % (r, n, p, "a store in flight" if st["auxwriting"][r][p]
                                       else "stored exchange %d" % st["auxtag"][r][p]))
            st["auxreading"][r] = 1 if kind == "aux_read_begin" else 0
        elif kind == "rec_swept":
            st["swept"][r] = n
        elif kind == "rec_pushed":
            st["pushed"][r][n % 2] = n

    seen = set()
    stack = [init()]
    while stack:
        st = stack.pop()
        k = key(st)
        if k in seen:
            continue
        seen.add(k)
        moved = False
        for s, (r, prog) in enumerate(streams):
            pc = st["pc"][s]
            if pc >= len(prog) or not enabled(st, r, prog[pc]):
                continue
            moved = True
            nx = clone(st)
            apply(nx, r, prog[pc])
            nx["pc"][s] = pc + 1
            stack.append(nx)
        if not moved and any(st["pc"][s] < len(streams[s][1]) for s in range(S)):
            raise Violation("deadlock at " + str([(streams[s][0], streams[s][1][st["pc"][s]])
                                                   for s in range(S) if st["pc"][s] < len(streams[s][1])]))
    return len(seen)


PAIR = [[1], [0]]
RING3 = [[1, 2], [0, 2], [0, 1]]                      # three ranks, everybody a peer of everybody
CHAIN3 = [[1], [0, 2], [1]]                           # slabs: the middle rank has two peers
ONE_WAY = [[1], []]                                   # rank 0 stores to rank 1, nothing comes back


@pytest.mark.parametrize("mode", ["wait-kernel", "sweep-wait", "overlap"])
@pytest.mark.parametrize("name,send,K", [("pair", PAIR, 5), ("chain3", CHAIN3, 3), ("ring3", RING3, 3)])
def test_state_exchange_needs_no_handshake_with_symmetric_peers(mode, name, send, K):
    if mode == "overlap" and name != "pair":
        pytest.skip("two streams per rank on three ranks: 20 s of state space for the non-default path; the pair covers it")
    n = explore(len(send), K, send, mode)
    assert n > 50


@pytest.mark.parametrize("mode", ["wait-kernel", "sweep-wait", "overlap"])
def test_one_directional_peers_race_which_is_why_connect_rejects_them(mode):
    with pytest.raises(Violation, match="reading|reads exchange"):
        explore(2, 4, ONE_WAY, mode)


@pytest.mark.parametrize("name,send", [("pair", PAIR), ("chain3", CHAIN3)])
def test_single_buffered_auxfield_rows_need_the_ready_handshake(name, send):
    K = 4 if name == "pair" else 3
    assert explore(len(send), K, send, "wait-kernel", aux=True, handshake=True) > 50
    with pytest.raises(Violation, match="auxField"):
        explore(len(send), K, send, "wait-kernel", aux=True, handshake=False)


def test_overlapped_push_is_ordered_against_the_sweep_two_steps_later():
    """api.cu waits for evPushed[parity] before the sweep that writes the buffer the push of two
    steps earlier read; without that event the model finds the overwrite"""
    assert explore(2, 4, PAIR, "overlap") > 50
    with pytest.raises(Violation, match="overwrites the buffer a push still reads|reads a buffer a sweep is writing"):
        explore(2, 4, PAIR, "overlap", no_evpushed=True)
