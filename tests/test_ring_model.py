"""A Python model of the producer's byte-ring bookkeeping in bucket_mul_v4_kernel (csrc/bucket_mul_v4.cuh, bulk
producer): units of 1..16 rows are placed first-in-first-out in a 16 KB ring, a unit never straddles the ring's end
(the bytes skipped at the wrap are charged to it), at most 16 units are outstanding, and the consumer's releases hand
the bytes back in order.  Checked: live units never overlap, accounting returns to "all free", empty records take no
slot."""
import random

RING, UNITS = 16 * 1024, 16


class Producer:
    def __init__(self):
        self.seq = self.tail = self.head = 0
        self.free = RING
        self.charged = [0] * UNITS       # lane s remembers what slot s holds
        self.live = {}                   # seq -> (off, bytes)   (model only)

    def reclaim(self, released_upto):
        while self.tail < released_upto:                       # in order: the consumer releases units first-in-first-out
            self.free += self.charged[self.tail % UNITS]
            del self.live[self.tail]
            self.tail += 1

    def place_window(self, sizes):
        """the allocation loop over one window of records; returns how many records were consumed"""
        h, f, n_ok, n_seen = self.head, self.free, 0, 0
        slots_free = UNITS - (self.seq - self.tail)
        for b in sizes:
            if b == 0:
                n_seen += 1
                continue
            skip = RING - h if h + b > RING else 0
            if f < b + skip or n_ok >= slots_free:
                break
            off = 0 if skip else h
            self.charged[(self.seq + n_ok) % UNITS] = b + skip
            self.live[self.seq + n_ok] = (off, b)
            h = off + b
            if h >= RING:
                h = 0
            f -= b + skip
            n_ok += 1
            n_seen += 1
        self.seq += n_ok
        self.head, self.free = h, f
        return n_seen


def _overlap(a, b):
    return a[0] < b[0] + b[1] and b[0] < a[0] + a[1]


def test_ring_never_overlaps_and_accounts_for_every_byte():
    rng = random.Random(7)
    for seg_bytes in (256, 64, 96):
        p = Producer()
        released = 0
        records = [rng.choice([0, 0, 1, 1, 2, 3, 5, 8, 13, 16]) * seg_bytes for _ in range(3000)]
        pos = 0
        while pos < len(records):
            window = records[pos:pos + rng.randint(1, 8)]
            done = 0
            while done < len(window):
                took = p.place_window(window[done:])
                live = list(p.live.values())
                for i in range(len(live)):
                    assert live[i][0] + live[i][1] <= RING                 # never straddles the end
                    for j in range(i):
                        assert not _overlap(live[i], live[j])
                assert p.seq - p.tail <= UNITS
                assert p.free == RING - sum(p.charged[s % UNITS] for s in range(p.tail, p.seq))
                if took == 0:                                              # blocked: the consumer releases the oldest unit
                    assert p.seq > p.tail, "a unit that does not fit an EMPTY ring would deadlock the pair"
                    released = max(released, p.tail) + 1
                    p.reclaim(released)
                else:
                    done += took
                    if rng.random() < 0.5 and p.seq > released:            # consumer progress, at its own pace
                        released += rng.randint(1, p.seq - released)
                        p.reclaim(released)
            pos += len(window)
        p.reclaim(p.seq)
        assert p.free == RING and not p.live
