"""Test doubles (TEST INFRASTRUCTURE): an in-memory AMQP broker and an Engine-shaped
wrapper around the CPU oracle, so host logic (search_worker.py, shard routing) is
covered on machines without a GPU.  Never imported by the product package."""
import collections
import importlib

import numpy as np


class FakeChannel:
    def __init__(self, broker, name):
        self.broker, self.name = broker, name
        self.unacked = {}
        self.prefetch = 0

    def basic_consume(self, queue, consumer):
        self.broker.consumers[queue].append((self, consumer))
        return f"ctag-{queue}-{len(self.broker.consumers[queue])}"

    def basic_publish(self, exchange, routing_key, payload, **props):
        self.broker.publish(exchange, routing_key, payload, props)

    def basic_ack(self, tag):
        self.unacked.pop(tag)
        self.broker.acked.append(tag)

    def basic_nack(self, tag):
        self.unacked.pop(tag)
        self.broker.nacked.append(tag)

    def queue_status(self, queue):
        return {"queue": queue, "message_count": len(self.broker.queues[queue]),
                "consumer_count": len(self.broker.consumers[queue])}


class FakeBroker:
    """Direct exchanges only (all the reference uses)."""

    def __init__(self):
        self.exchanges = {}
        self.queues = collections.defaultdict(collections.deque)
        self.bindings = collections.defaultdict(list)  # (exchange, routing_key) -> [queue]
        self.consumers = collections.defaultdict(list)
        self.acked, self.nacked = [], []
        self._tag = 0

    # connection API used by SearchWorker.start_link
    def spawn_channel(self, name):
        return FakeChannel(self, name)

    def configure_channel(self, channel, config):
        ex, q = config["exchange"], config["queue"]
        self.exchanges[ex["name"]] = ex
        self.queues[q["name"]]
        self.bindings[(ex["name"], q["routing_key"])].append(q["name"])
        channel.prefetch = config["qos"]["prefetch_count"]

    def bind(self, exchange, routing_key, queue):
        self.exchanges.setdefault(exchange, {"name": exchange, "type": "direct"})
        self.queues[queue]
        self.bindings[(exchange, routing_key)].append(queue)

    def publish(self, exchange, routing_key, payload, props=None):
        for q in self.bindings.get((exchange, routing_key), []):
            self.queues[q].append((payload, props or {}))

    def deliver_all(self):
        """Push queued messages to their consumers (round-robin), like the broker would: a channel with
        `prefetch` unacknowledged deliveries (basic.qos, search/worker.ex:29) gets nothing more until it acks."""
        n = 0
        for q, cons in list(self.consumers.items()):
            i = 0
            while self.queues[q] and cons:
                ready = [(ch, c) for ch, c in cons if not ch.prefetch or len(ch.unacked) < ch.prefetch]
                if not ready:
                    break  # every consumer of this queue is at its QoS limit
                ch, consumer = ready[i % len(ready)]
                i += 1
                payload, props = self.queues[q].popleft()
                self._tag += 1
                ch.unacked[self._tag] = payload
                consumer.handle_info(("basic_deliver", payload, {"delivery_tag": self._tag, **props}))
                n += 1
        return n


class OracleEngine:
    """Engine API on top of the CPU oracle — for CPU-only tests of host logic."""

    def __init__(self, cfg):
        self.cfg = cfg
        self.oracle = importlib.import_module("oracle.oracle")
        self.abi = importlib.import_module("microservice-matchmaking_b200.abi")
        self.q = [np.zeros(0, np.uint64), np.zeros(0, np.int32), np.zeros(0, np.uint8)]
        self.alive = np.zeros(0, np.uint8)
        self.active = set()
        self.max_spread = -1
        self.dead_pending = 0

    def set_option(self, name, value):
        if name != "max_spread":
            raise ValueError(name)
        if value >= 0 and self.cfg.order_mode != self.abi.MM_ORDER_RATING:
            raise ValueError("max_spread needs MM_ORDER_RATING")
        self.max_spread = int(value)

    def enqueue(self, ids, rating, mode, enq_ts=None):
        ids = np.asarray(ids, np.uint64); rating = np.asarray(rating, np.int32); mode = np.asarray(mode, np.uint8)
        acc = np.zeros(len(ids), np.uint8)
        keep = []
        for i, (p, r, m) in enumerate(zip(ids.tolist(), rating.tolist(), mode.tolist())):
            if m >= self.cfg.n_modes or p >= 2 ** 64 - 2 or self.oracle.find_rating_group(self.cfg, r) < 0:
                acc[i] = 2
            elif p in self.active:
                acc[i] = 0
            elif len(self.q[0]) + len(keep) >= self.cfg.capacity:
                acc[i] = 3
            else:
                acc[i] = 1
                self.active.add(p)
                keep.append(i)
        self.q = [np.concatenate([a, b[keep]]) for a, b in zip(self.q, (ids, rating, mode))]
        self.alive = np.concatenate([self.alive, np.ones(len(keep), np.uint8)])
        return acc

    def remove(self, ids):
        """Like the GPU engine, the queued entry of a leaver is dropped for good (the engine tombstones it): an id that
        is removed and enqueued again — e.g. a recycled host handle — is a NEW entry.  (In the serialized reference the
        stale entry would pass in_queue? again and show up twice; DESIGN.md §2 lists it as a reference defect.)"""
        n = 0
        for p in np.asarray(ids, np.uint64).tolist():
            if p in self.active:
                self.active.discard(p)
                keep = self.q[0] != p
                self.dead_pending += int((~keep).sum())
                self.q = [a[keep] for a in self.q]
                self.alive = self.alive[keep]
                n += 1
        return n

    def take(self, ids):
        """mm_take: leave the pool, stay in the active set."""
        ids = np.asarray(ids, np.uint64)
        hit = np.isin(self.q[0], ids)
        self.q = [a[~hit] for a in self.q]
        self.alive = self.alive[~hit]
        return int(hit.sum())

    def in_queue(self, ids):
        return np.array([p in self.active for p in np.asarray(ids, np.uint64).tolist()], bool)

    def pool_size(self):
        return len(self.q[0])

    def tick(self, now=0):
        if self.max_spread >= 0:
            ref = self.oracle.run_windowed(self.cfg, self.max_spread, *self.q, alive=self.alive)
        else:
            ref = self.oracle.run_literal(self.cfg, *self.q, alive=self.alive)
        keep = np.isin(self.q[0], ref.residual_ids)
        self.q = [a[keep] for a in self.q]
        self.alive = np.ones(len(self.q[0]), np.uint8)
        st = self.abi.TickStats()
        st.n_lobbies, st.n_matched, st.n_residual, st.n_dead = ref.n_lobbies, ref.n_matched, ref.n_residual, ref.n_dead + self.dead_pending
        self.dead_pending = 0
        return ref.lobbies, ref.member_ids, ref.emit_seq, st

    def pool_read(self):
        k = self.alive.astype(bool)
        return dict(id=self.q[0][k], rating=self.q[1][k], mode=self.q[2][k])

    def status(self):
        return {"message_count": self.pool_size(), "active_count": len(self.active)}

    def close(self):
        pass
