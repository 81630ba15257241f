// nexmark_gen.cpp — seeded, column-major Nexmark event generator (host side).
//
// Workload driver, not a kernel: it restates the *distributions* of the
// reference generator (crates/nexmark/src/generator/{mod,bids,auctions,people,
// price,config}.rs) with a counter-based RNG, because the reference seeds
// from ThreadRng and is not reproducible (crates/nexmark/src/lib.rs:198).
// Only the columns q3/q4/q7 read are produced.  Event i is a
// Person / Auction / Bid by i % 50 (generator/mod.rs:77-88, proportions
// 1:3:46 config.rs:26,46,91).
//
// Strings are carried as order-preserving dictionary codes: state = rank in
// the (already sorted) US_STATES list (people.rs:18-25), city = rank of the
// city name among the 10 sorted names (people.rs:27-38), name = rank of
// "First Last" among the 99 sorted combinations (people.rs:40-46); a bid's
// random filler `extra` (bids.rs:113) is a 32-bit handle.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <string>
#include <thread>
#include <vector>

typedef uint64_t u64;

namespace {
const u64 PERSON_PROP = 1, AUCTION_PROP = 3, BID_PROP = 46, TOTAL_PROP = 50;   // config.rs:26,46,91
const u64 FIRST_PERSON_ID = 1000, FIRST_AUCTION_ID = 1000, FIRST_CATEGORY_ID = 10;  // generator/config.rs:5-7
const u64 NUM_CATEGORIES = 5;          // auctions.rs:19
const u64 HOT_AUCTION_RATIO = 2, HOT_BIDDERS_RATIO = 4, HOT_SELLERS_RATIO = 4;  // config.rs:54-63
const u64 HOT_RATIO_DIV = 100;         // bids.rs:12-13, auctions.rs:21
const u64 NUM_ACTIVE_PEOPLE = 1000, PERSON_ID_LEAD = 10, NUM_IN_FLIGHT_AUCTIONS = 100;  // config.rs:11,71,81
const u64 BASE_TIME = 1436918400000ull;   // 2015-07-15T00:00:00Z, the Nexmark epoch
const u64 DEFAULT_FIRST_EVENT_RATE = 10000000;   // NexmarkConfig::first_event_rate (config.rs:51,134), a CLI option there

inline u64 mix64(u64 x) {
  x += 0x9e3779b97f4a7c15ull;
  x = (x ^ (x >> 30)) * 0xbf58476d1ce4e5b9ull;
  x = (x ^ (x >> 27)) * 0x94d049bb133111ebull;
  return x ^ (x >> 31);
}
struct Rng {   // counter-based: stream (seed, event, draw#)
  u64 s, e, k;
  Rng(u64 seed, u64 event) : s(seed), e(event), k(0) {}
  u64 next() { return mix64(mix64(s ^ (e * 0x9e3779b97f4a7c15ull)) + (k++) * 0xd1342543de82ef95ull); }
  u64 range(u64 n) { return (u64)(((unsigned __int128)next() * n) >> 64); }   // uniform in [0,n)
  float unit() { return (float)(next() >> 40) * (1.0f / 16777216.0f); }      // [0,1)
};

// generator/config.rs:118-120
inline u64 timestamp_for_event(double delay_us, u64 n) { return BASE_TIME + (u64)(delay_us * (double)n) / 1000; }
// people.rs:105-113
inline u64 last_base0_person_id(u64 event_id) {
  u64 epoch = event_id / TOTAL_PROP, offset = event_id % TOTAL_PROP;
  if (offset >= PERSON_PROP) offset = PERSON_PROP - 1;
  return epoch * PERSON_PROP + offset;
}
// people.rs:93-103
inline u64 next_base0_person_id(Rng& r, u64 event_id) {
  u64 num_people = last_base0_person_id(event_id) + 1;
  u64 active = std::min(num_people, NUM_ACTIVE_PEOPLE);
  u64 n = r.range(active + PERSON_ID_LEAD);
  return num_people - active + n;
}
// auctions.rs:85-108
inline u64 last_base0_auction_id(u64 event_id) {
  u64 epoch = event_id / TOTAL_PROP, offset = event_id % TOTAL_PROP;
  if (offset < PERSON_PROP) {
    if (epoch == 0) return 0;
    epoch -= 1;
    offset = AUCTION_PROP - 1;
  } else if (offset >= PERSON_PROP + AUCTION_PROP) {
    offset = AUCTION_PROP - 1;
  } else {
    offset -= PERSON_PROP;
  }
  return epoch * AUCTION_PROP + offset;
}
// auctions.rs:110-117
inline u64 next_base0_auction_id(Rng& r, u64 event_id) {
  u64 maxa = last_base0_auction_id(event_id);
  u64 mina = maxa > NUM_IN_FLIGHT_AUCTIONS ? maxa - NUM_IN_FLIGHT_AUCTIONS : 0;
  return mina + r.range(maxa - mina + 1);
}
// price.rs:9-11
inline u64 next_price(Rng& r) { return (u64)std::ceil(std::pow(10.0f, r.unit() * 6.0f) * 100.0f); }

struct Dict {
  u64 name_rank[11 * 9], city_rank[10];
  Dict() {
    const char* first[11] = {"Peter", "Paul", "Luke", "John", "Saul", "Vicky", "Kate", "Julie", "Sarah", "Deiter", "Walter"};
    const char* last[9] = {"Shultz", "Abrams", "Spencer", "White", "Bartels", "Walton", "Smith", "Jones", "Noris"};
    const char* city[10] = {"Phoenix", "Los Angeles", "San Francisco", "Boise", "Portland", "Bend", "Redmond", "Seattle", "Kent", "Cheyenne"};
    std::vector<std::pair<std::string, int>> v;
    for (int f = 0; f < 11; f++) for (int l = 0; l < 9; l++) v.push_back({std::string(first[f]) + " " + last[l], f * 9 + l});
    std::sort(v.begin(), v.end());
    for (size_t i = 0; i < v.size(); i++) name_rank[v[i].second] = i;
    std::vector<std::pair<std::string, int>> c;
    for (int i = 0; i < 10; i++) c.push_back({city[i], i});
    std::sort(c.begin(), c.end());
    for (size_t i = 0; i < c.size(); i++) city_rank[c[i].second] = i;
  }
};
const Dict& dict() { static Dict d; return d; }

inline u64 persons_before(u64 e) { return (e / TOTAL_PROP) * PERSON_PROP + std::min(e % TOTAL_PROP, PERSON_PROP); }
inline u64 auctions_before(u64 e) {
  u64 r = e % TOTAL_PROP;
  u64 a = r <= PERSON_PROP ? 0 : std::min(r - PERSON_PROP, AUCTION_PROP);
  return (e / TOTAL_PROP) * AUCTION_PROP + a;
}
inline u64 bids_before(u64 e) {
  u64 r = e % TOTAL_PROP;
  u64 b = r <= PERSON_PROP + AUCTION_PROP ? 0 : r - PERSON_PROP - AUCTION_PROP;
  return (e / TOTAL_PROP) * BID_PROP + b;
}

struct Cols {
  u64 *p_id, *p_name, *p_city, *p_state, *p_dt;
  u64 *a_id, *a_seller, *a_category, *a_dt, *a_expires;
  u64 *b_auction, *b_bidder, *b_price, *b_dt, *b_extra;
};

void gen_range(u64 seed, double delay_us, u64 first, u64 lo, u64 hi, const Cols& c) {
  const Dict& d = dict();
  u64 p0 = persons_before(first), a0 = auctions_before(first), b0 = bids_before(first);
  for (u64 e = lo; e < hi; e++) {
    Rng r(seed, e);
    u64 rem = e % TOTAL_PROP;
    u64 ts = timestamp_for_event(delay_us, e);
    if (rem < PERSON_PROP) {   // people.rs:50-86
      u64 i = persons_before(e) - p0;
      if (c.p_id) c.p_id[i] = last_base0_person_id(e) + FIRST_PERSON_ID;
      u64 f = r.range(11), l = r.range(9), city = r.range(10), state = r.range(6);
      if (c.p_name) c.p_name[i] = d.name_rank[f * 9 + l];
      if (c.p_city) c.p_city[i] = d.city_rank[city];
      if (c.p_state) c.p_state[i] = state;
      if (c.p_dt) c.p_dt[i] = ts;
    } else if (rem < PERSON_PROP + AUCTION_PROP) {   // auctions.rs:27-83
      u64 i = auctions_before(e) - a0;
      if (c.a_id) c.a_id[i] = last_base0_auction_id(e) + FIRST_AUCTION_ID;
      u64 seller = (r.range(HOT_SELLERS_RATIO) == 0)
                       ? next_base0_person_id(r, e)
                       : (last_base0_person_id(e) / HOT_RATIO_DIV) * HOT_RATIO_DIV;
      if (c.a_seller) c.a_seller[i] = seller + FIRST_PERSON_ID;
      if (c.a_category) c.a_category[i] = FIRST_CATEGORY_ID + r.range(NUM_CATEGORIES);
      if (c.a_dt) c.a_dt[i] = ts;
      // auctions.rs:124-143 next_auction_length_ms
      u64 num_events_for_auctions = (NUM_IN_FLIGHT_AUCTIONS * TOTAL_PROP) / AUCTION_PROP;
      u64 future = timestamp_for_event(delay_us, e + num_events_for_auctions);
      u64 horizon = future > ts ? future - ts : 0;
      u64 len = 1 + r.range(std::max<u64>(horizon * 2, 1));
      if (c.a_expires) c.a_expires[i] = ts + len;
    } else {   // bids.rs:60-115
      if (!c.b_auction && !c.b_bidder && !c.b_price && !c.b_dt && !c.b_extra) {   // table not requested
        e += (TOTAL_PROP - 1 - rem);   // skip to the end of this 50-event epoch
        continue;
      }
      u64 i = bids_before(e) - b0;
      u64 auction = (r.range(HOT_AUCTION_RATIO) == 0)
                        ? next_base0_auction_id(r, e)
                        : (last_base0_auction_id(e) / HOT_RATIO_DIV) * HOT_RATIO_DIV;
      u64 bidder = (r.range(HOT_BIDDERS_RATIO) == 0)
                       ? next_base0_person_id(r, e)
                       : (last_base0_person_id(e) / HOT_RATIO_DIV) * HOT_RATIO_DIV + 1;
      if (c.b_auction) c.b_auction[i] = auction + FIRST_AUCTION_ID;
      if (c.b_bidder) c.b_bidder[i] = bidder + FIRST_PERSON_ID;
      if (c.b_price) c.b_price[i] = next_price(r);
      if (c.b_dt) c.b_dt[i] = ts;
      if (c.b_extra) c.b_extra[i] = r.next() >> 32;
    }
  }
}
}  // namespace

extern "C" {

// Rows of each table contributed by events [first, first+n).
void nexmark_counts(u64 first, u64 n, u64* np, u64* na, u64* nb) {
  *np = persons_before(first + n) - persons_before(first);
  *na = auctions_before(first + n) - auctions_before(first);
  *nb = bids_before(first + n) - bids_before(first);
}

// Fill the requested columns (NULL = skip) for events [first, first+n).  `rate` = first_event_rate in
// events/s (generator/config.rs:62: inter_event_delay_us = 1e6 / rate); 0 = the reference default.
void nexmark_generate_rate(u64 seed, u64 rate, u64 first, u64 n, int nthreads, u64* p_id, u64* p_name, u64* p_city,
                           u64* p_state, u64* p_dt, u64* a_id, u64* a_seller, u64* a_category, u64* a_dt, u64* a_expires,
                           u64* b_auction, u64* b_bidder, u64* b_price, u64* b_dt, u64* b_extra) {
  Cols c{p_id, p_name, p_city, p_state, p_dt, a_id, a_seller, a_category, a_dt, a_expires, b_auction, b_bidder, b_price, b_dt, b_extra};
  const double delay_us = 1000000.0 / (double)(rate ? rate : DEFAULT_FIRST_EVENT_RATE);
  if (nthreads < 1) nthreads = 1;
  if (n < 100000) nthreads = 1;
  std::vector<std::thread> th;
  u64 chunk = (n + nthreads - 1) / nthreads;
  for (int t = 0; t < nthreads; t++) {
    u64 lo = first + std::min<u64>(n, t * chunk), hi = first + std::min<u64>(n, (t + 1) * chunk);
    if (lo >= hi) continue;
    th.emplace_back([=] { gen_range(seed, delay_us, first, lo, hi, c); });
  }
  for (auto& t : th) t.join();
}
void nexmark_generate(u64 seed, u64 first, u64 n, int nthreads, u64* p_id, u64* p_name, u64* p_city, u64* p_state,
                      u64* p_dt, u64* a_id, u64* a_seller, u64* a_category, u64* a_dt, u64* a_expires, u64* b_auction,
                      u64* b_bidder, u64* b_price, u64* b_dt, u64* b_extra) {
  nexmark_generate_rate(seed, 0, first, n, nthreads, p_id, p_name, p_city, p_state, p_dt, a_id, a_seller, a_category, a_dt,
                        a_expires, b_auction, b_bidder, b_price, b_dt, b_extra);
}

u64 nexmark_base_time(void) { return BASE_TIME; }

}  // extern "C"
