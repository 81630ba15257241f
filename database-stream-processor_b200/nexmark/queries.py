"""Nexmark q0 / q3 / q4 / q7 on the Stream API.

The reference's `NexmarkStream` is one stream of `Event` enum rows
(crates/nexmark/src/queries/mod.rs:12); here it is three column tables
(person / auction / bid) per step, which is the same data column-major.
Column order: see generator.PERSON_COLS / AUCTION_COLS / BID_COLS.
"""
from __future__ import annotations

from dataclasses import dataclass

from ..circuit import Max, Min, RootCircuit, Stream, TableStream
from ..zset import Proj, Schema, col, key, lval, rval

WATERMARK_INTERVAL_SECONDS = 4  # queries/mod.rs:12
TUMBLE_SECONDS = 10             # queries/q7.rs:43
CATEGORY_OF_INTEREST = 10       # queries/q3.rs:31


@dataclass
class NexmarkTables:
    person: TableStream
    auction: TableStream
    bid: TableStream


def add_nexmark_input(circuit: RootCircuit):
    p, hp = circuit.add_input_table(5)
    a, ha = circuit.add_input_table(5)
    b, hb = circuit.add_input_table(5)
    return NexmarkTables(p, a, b), {"person": hp, "auction": ha, "bid": hb}


def q0(inp: NexmarkTables) -> Stream:
    """q0 pass-through of bids (queries/q0.rs): plumbing only."""
    return inp.bid.flat_map_index(Proj(Schema("uuuuu"), [col(0), col(1), col(2), col(3), col(4)]))


def q3(inp: NexmarkTables, states_of_interest=(1, 2, 3)) -> Stream:
    """queries/q3.rs:35-63.  States of interest OR, ID, CA = codes 3, 2, 1."""
    # Select auctions of interest and index them by seller id (q3.rs:37-40).
    auction_by_seller = inp.auction.flat_map_index(
        Proj(Schema("u", "u"), [col(1), col(0)], where=[col(2).eq(CATEGORY_OF_INTEREST)]))
    # Select people from states of interest, indexed by person id (q3.rs:43-50).
    person_by_id = inp.person.flat_map_index(
        Proj(Schema("u", "uuu"), [col(0), col(1), col(2), col(3)], where=[col(3).isin(states_of_interest)]))
    # (name, city, state, auction_id) (q3.rs:53-63)
    return auction_by_seller.join(person_by_id, Proj(Schema("uuuu"), [rval(0), rval(1), rval(2), lval(0)]))


def q4(inp: NexmarkTables) -> Stream:
    """queries/q4.rs:43-83."""
    # (a.id, (a.category, a.date_time, a.expires))  (q4.rs:45-48)
    auctions_by_id = inp.auction.flat_map_index(Proj(Schema("u", "uuu"), [col(0), col(2), col(3), col(4)]))
    # (b.auction, (b.price, b.date_time))  (q4.rs:51-54)
    bids_by_auction = inp.bid.flat_map_index(Proj(Schema("u", "uu"), [col(0), col(2), col(3)]))
    # ((auction_id, category), bid_price) for bids inside [date_time, expires]  (q4.rs:58-67)
    bids_for_auctions = auctions_by_id.join_index(
        bids_by_auction,
        Proj(Schema("uu", "u"), [key(0), lval(0), rval(0)], where=[rval(1).ge(lval(1)), rval(1).le(lval(2))]))
    winning_bids = bids_for_auctions.aggregate(Max)                                    # q4.rs:73-74
    by_category = winning_bids.map_index(Proj(Schema("u", "u"), [key(1), lval(0)]))    # q4.rs:75-76
    avg = by_category.average(lval(0))                                                 # q4.rs:80-81
    return avg.map(Proj(Schema("uu"), [key(0), lval(0)]))                              # q4.rs:82


def q7(inp: NexmarkTables) -> Stream:
    """queries/q7.rs:45-94."""
    # (date_time, (auction, bidder, price, extra))  (q7.rs:47-55)
    bids_by_time = inp.bid.flat_map_index(Proj(Schema("u", "uuuu"), [col(3), col(0), col(1), col(2), col(4)]))
    watermark = bids_by_time.watermark_monotonic(lambda dt: dt - WATERMARK_INTERVAL_SECONDS * 1000)   # q7.rs:60-61

    def bounds(wm):   # q7.rs:64-70
        rounded = wm - (wm % (TUMBLE_SECONDS * 1000))
        return (max(rounded - TUMBLE_SECONDS * 1000, 0), rounded)

    windowed = bids_by_time.window(watermark.apply(bounds))                            # q7.rs:73
    # (price, (auction, bidder, price, date_time, extra))  (q7.rs:74-79)
    bids_by_price = windowed.map_index(Proj(Schema("u", "uuuuu"), [lval(2), lval(0), lval(1), lval(2), key(0), lval(3)]))
    # ((), -price) -> Min -> max price  (q7.rs:82-91)
    neg_price = windowed.map_index(Proj(Schema("", "i"), [-lval(2)]))
    if windowed.circuit.workers > 1:
        # The single group () would shard every windowed bid to one worker.  Min is a
        # semigroup (aggregate/min.rs:17-27): aggregate per worker first, then
        # aggregate the <= P partial minima.
        neg_price = neg_price.aggregate(Min, local=True)
    max_price = neg_price.aggregate(Min).map(Proj(Schema("u"), [-lval(0)]))
    # all bids with the max price  (q7.rs:92-93)
    return max_price.join(bids_by_price, Proj(Schema("uuuuu"), [rval(0), rval(1), rval(2), rval(3), rval(4)]))


QUERIES = {"q0": q0, "q3": q3, "q4": q4, "q7": q7}
