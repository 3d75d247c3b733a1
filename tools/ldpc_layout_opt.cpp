// tools/ldpc_layout_opt.cpp -- offline optimiser of the decoder's shared-memory layout: mercury_b200/data/ldpc_tables.bin -> ldpc_layout.bin ('MLAY').
//
//   g++ -O2 -std=c++17 tools/ldpc_layout_opt.cpp -o /tmp/ldpc_layout_opt && /tmp/ldpc_layout_opt mercury_b200/data/ldpc_tables.bin mercury_b200/data/ldpc_layout.bin [moves]
//
// mb_ldpc_kernel keeps posterior[1600] and one message per Tanner-graph edge in shared memory.  A warp owns 32 checks (variables)
// and at step k every lane gathers through its k-th edge: lane i of a check group reads posterior[variable(i, k)], lane i of a
// variable group reads message[slot of edge (i, k)].  The bank of a posterior is the variable's position inside ITS group of 32, the
// bank of a message is the check's position inside ITS group of 32, so a random graph costs ~3 shared-memory wavefronts per gather.
// Two things are free without touching the kernel or the graph: the ORDER of a node's edges, and the order of the nodes INSIDE a
// group of 32 (membership of the groups -- nodes sorted by degree -- stays as it is, so the two sides decouple: check-side gathers
// depend on the checks' edge orders and the variables' in-group positions, variable-side gathers on the variables' edge orders and
// the checks' in-group positions).  This tool minimises sum over (group, step) of the largest bank multiplicity by hill climbing with
// plateau moves (fixed seed) and writes the layout file: new node orders and edge orders.  The decoder's arithmetic is unchanged
// up to the order in which fp32 sums run; mb_tables.cpp validates the file (permutations of the reference rows) and uses the
// reference order without it.  The tables file itself (what tools/extract_ldpc_tables.py derives from the reference) is not touched.
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <string>
#include <vector>

struct Rate {
	uint16_t h[6];  // rate_num, N, K, P, Cwidth, Vwidth
	uint32_t ne;
	std::vector<std::vector<int>> crow, vrow;
};

static uint64_t rng_state = 0x9E3779B97F4A7C15ull;
static inline uint32_t rnd()
{
	rng_state ^= rng_state << 13, rng_state ^= rng_state >> 7, rng_state ^= rng_state << 17;
	return (uint32_t)(rng_state >> 32);
}

// One side of the problem.  `rows[n]` = targets of gathering node n (order is optimised); `groups` = the gathering nodes in warp
// groups of 32 (fixed); `tgroup[t]`, `tlane[t]` = group and in-group position of target t (lane is optimised by swaps inside a group).
struct Side {
	std::vector<std::vector<int>> *rows;
	std::vector<std::vector<int>> groups;
	std::vector<int> node_group, tgroup, tlane, tdeg;  // tdeg: the target's own degree -- only equal-degree targets trade places (groups stay sorted)
	std::vector<std::vector<int>> tmembers;           // targets of each target group
	std::vector<std::vector<int>> cost;               // [group][step]
	bool broadcast = true;  // two lanes naming the same target read the same word (a posterior: yes; messages of one check: no, same bank, different words)

	// objective of the search: 64 * (largest bank load) + sum of squared bank loads -- the first term is what the hardware pays (wavefronts),
	// the second gives the climb a slope across the plateaus where the maximum does not change yet
	int step_cost(int g, int k, bool wavefronts_only = false) const
	{
		int cnt[32] = {0}, m = 1;
		int seen[32], ns = 0;
		bool had_pad = false;
		for (int n : groups[g]) {
			const std::vector<int> &r = (*rows)[n];
			if (k >= (int)r.size()) {  // padding: every padded lane reads the one neutral word, which sits in the bank of position 0
				if (!had_pad) had_pad = true, m = std::max(m, ++cnt[0]);
				continue;
			}
			const int t = r[k];
			bool dup = false;  // the same word read twice is a broadcast, not a conflict
			for (int i = 0; broadcast && i < ns; i++) dup |= seen[i] == t;
			if (dup) continue;
			if (broadcast) seen[ns++] = t;
			m = std::max(m, ++cnt[tlane[t]]);
		}
		if (wavefronts_only) return m;
		int sq = 0;
		for (int b = 0; b < 32; b++) sq += cnt[b] * cnt[b];
		return 64 * m + sq;
	}
	long wavefronts() const
	{
		long s = 0;
		for (size_t g = 0; g < cost.size(); g++)
			for (size_t k = 0; k < cost[g].size(); k++) s += step_cost((int)g, (int)k, true);
		return s;
	}
	long total() const
	{
		long s = 0;
		for (auto &g : cost)
			for (int c : g) s += c;
		return s;
	}
	void init()
	{
		cost.assign(groups.size(), {});
		for (size_t g = 0; g < groups.size(); g++) {
			size_t d = 0;
			for (int n : groups[g]) d = std::max(d, (*rows)[n].size());
			cost[g].resize(d);
			for (size_t k = 0; k < d; k++) cost[g][k] = step_cost((int)g, (int)k);
		}
	}
	void run(long moves)
	{
		std::vector<int> nodes;
		for (size_t n = 0; n < rows->size(); n++)
			if ((*rows)[n].size() >= 2) nodes.push_back((int)n);
		std::vector<int> tg;
		for (size_t g = 0; g < tmembers.size(); g++)
			if (tmembers[g].size() >= 2) tg.push_back((int)g);
		// which (group, step) pairs a target appears in: positions change with edge swaps, so look them up through the rows
		std::vector<std::vector<int>> users(tgroup.size());  // target -> gathering nodes
		for (size_t n = 0; n < rows->size(); n++)
			for (int t : (*rows)[n]) users[t].push_back((int)n);
		for (long it = 0; it < moves; it++) {
			if (rnd() & 1) {  // swap two edges of one node
				const int n = nodes[rnd() % nodes.size()];
				std::vector<int> &r = (*rows)[n];
				const int i = rnd() % r.size();
				int j = rnd() % (r.size() - 1);
				if (j >= i) j++;
				const int g = node_group[n];
				const int old = cost[g][i] + cost[g][j];
				std::swap(r[i], r[j]);
				const int ci = step_cost(g, i), cj = step_cost(g, j);
				if (ci + cj <= old) cost[g][i] = ci, cost[g][j] = cj;
				else std::swap(r[i], r[j]);
			} else {  // swap the in-group positions of two targets of one target group
				const std::vector<int> &mem = tmembers[tg[rnd() % tg.size()]];
				const int a = mem[rnd() % mem.size()];
				int bi = rnd() % (mem.size() - 1);
				const int b = mem[bi] == a ? mem[mem.size() - 1] : mem[bi];
				if (a == b || tdeg[a] != tdeg[b]) continue;
				std::pair<int, int> aff[128];
				int na = 0;
				for (int t : {a, b})
					for (int n : users[t]) {
						const std::vector<int> &r = (*rows)[n];
						const int k = (int)(std::find(r.begin(), r.end(), t) - r.begin());
						const std::pair<int, int> p(node_group[n], k);
						bool dup = false;
						for (int q = 0; q < na; q++) dup |= aff[q] == p;
						if (!dup && na < 128) aff[na++] = p;
					}
				int old = 0, neu = 0, nc[128];
				for (int q = 0; q < na; q++) old += cost[aff[q].first][aff[q].second];
				std::swap(tlane[a], tlane[b]);
				for (int q = 0; q < na; q++) neu += nc[q] = step_cost(aff[q].first, aff[q].second);
				if (neu <= old)
					for (int q = 0; q < na; q++) cost[aff[q].first][aff[q].second] = nc[q];
				else std::swap(tlane[a], tlane[b]);
			}
		}
	}
};

int main(int argc, char **argv)
{
	if (argc < 3) return fprintf(stderr, "usage: %s ldpc_tables.bin ldpc_layout.bin [moves per side]\n", argv[0]), 2;
	const long moves = argc > 3 ? atol(argv[3]) : 4000000;
	FILE *f = fopen(argv[1], "rb");
	if (!f) return perror(argv[1]), 1;
	std::vector<uint8_t> file;
	{
		uint8_t buf[65536];
		size_t n;
		while ((n = fread(buf, 1, sizeof(buf), f)) > 0) file.insert(file.end(), buf, buf + n);
		fclose(f);
	}
	if (file.size() < 12 || memcmp(file.data(), "MLDP", 4) != 0) return fprintf(stderr, "bad header\n"), 1;
	uint32_t nr;
	memcpy(&nr, file.data() + 8, 4);
	size_t off = 12;
	std::vector<Rate> rates(nr);
	for (Rate &r : rates) {
		memcpy(r.h, file.data() + off, 12), off += 12;
		memcpy(&r.ne, file.data() + off, 4), off += 4;
		const int N = r.h[1], P = r.h[3];
		const uint16_t *cdeg = (const uint16_t *)(file.data() + off);
		off += 2 * P;
		const uint16_t *ev = (const uint16_t *)(file.data() + off);
		off += 2 * r.ne;
		const uint16_t *vdeg = (const uint16_t *)(file.data() + off);
		off += 2 * N;
		const uint16_t *vc = (const uint16_t *)(file.data() + off);
		off += 2 * r.ne;
		size_t e = 0;
		r.crow.resize(P);
		for (int c = 0; c < P; c++)
			for (int j = 0; j < cdeg[c]; j++) r.crow[c].push_back(ev[e++]);
		e = 0;
		r.vrow.resize(N);
		for (int v = 0; v < N; v++)
			for (int j = 0; j < vdeg[v]; j++) r.vrow[v].push_back(vc[e++]);
	}
	std::vector<uint8_t> lay;
	auto put16 = [&](uint16_t v) { lay.push_back((uint8_t)(v & 0xFF)), lay.push_back((uint8_t)(v >> 8)); };
	auto put32 = [&](uint32_t v) { put16((uint16_t)(v & 0xFFFF)), put16((uint16_t)(v >> 16)); };
	lay.insert(lay.end(), {'M', 'L', 'A', 'Y'});
	put32(1), put32(nr);
	for (Rate &r : rates) {
		const int N = r.h[1], P = r.h[3];
		std::vector<int> csorted(P), vsorted(N);
		std::iota(csorted.begin(), csorted.end(), 0);
		std::stable_sort(csorted.begin(), csorted.end(), [&](int a, int b) { return r.crow[a].size() > r.crow[b].size(); });
		std::iota(vsorted.begin(), vsorted.end(), 0);
		std::stable_sort(vsorted.begin(), vsorted.end(), [&](int a, int b) { return r.vrow[a].size() > r.vrow[b].size(); });
		auto make_side = [&](std::vector<std::vector<int>> &rows, const std::vector<int> &gsorted, const std::vector<int> &tsorted,
				     const std::vector<std::vector<int>> &trows) {
			Side s;
			s.rows = &rows;
			s.node_group.assign(gsorted.size(), 0);
			for (size_t i = 0; i < gsorted.size(); i++) {
				if (i % 32 == 0) s.groups.emplace_back();
				s.groups.back().push_back(gsorted[i]);
				s.node_group[gsorted[i]] = (int)(i / 32);
			}
			s.tgroup.assign(tsorted.size(), 0), s.tlane.assign(tsorted.size(), 0), s.tdeg.assign(tsorted.size(), 0);
			for (size_t t = 0; t < tsorted.size(); t++) s.tdeg[t] = (int)trows[t].size();
			s.tmembers.assign((tsorted.size() + 31) / 32, {});
			for (size_t i = 0; i < tsorted.size(); i++) {
				s.tgroup[tsorted[i]] = (int)(i / 32), s.tlane[tsorted[i]] = (int)(i % 32);
				s.tmembers[i / 32].push_back(tsorted[i]);
			}
			s.init();
			return s;
		};
		Side cs = make_side(r.crow, csorted, vsorted, r.vrow);  // check-side gathers of posteriors
		Side vs = make_side(r.vrow, vsorted, csorted, r.crow);  // variable-side gathers of messages
		vs.broadcast = false;
		vs.init();
		const long c0 = cs.wavefronts(), v0 = vs.wavefronts();
		size_t csteps = 0, vsteps = 0;
		for (auto &g : cs.cost) csteps += g.size();
		for (auto &g : vs.cost) vsteps += g.size();
		cs.run(moves), vs.run(moves);
		printf("rate %2d/16: check-side gathers %ld -> %ld wavefronts over %zu steps (%.2f -> %.2f), variable-side %ld -> %ld over %zu (%.2f -> %.2f)\n", r.h[0], c0,
		       cs.wavefronts(), csteps, (double)c0 / csteps, (double)cs.wavefronts() / csteps, v0, vs.wavefronts(), vsteps, (double)v0 / vsteps,
		       (double)vs.wavefronts() / vsteps);
		// new node orders: group membership kept, in-group position = optimised lane
		std::vector<int> vnew(N), cnew(P);
		for (int v = 0; v < N; v++) vnew[cs.tgroup[v] * 32 + cs.tlane[v]] = v;
		for (int c = 0; c < P; c++) cnew[vs.tgroup[c] * 32 + vs.tlane[c]] = c;
		put16(r.h[0]), put16((uint16_t)N), put16((uint16_t)P), put16(0), put32(r.ne);
		for (int v : vnew) put16((uint16_t)v);
		for (int c : cnew) put16((uint16_t)c);
		for (int c = 0; c < P; c++)
			for (int v : r.crow[c]) put16((uint16_t)v);
		for (int v = 0; v < N; v++)
			for (int c : r.vrow[v]) put16((uint16_t)c);
	}
	f = fopen(argv[2], "wb");
	if (!f) return perror(argv[2]), 1;
	fwrite(lay.data(), 1, lay.size(), f);
	fclose(f);
	printf("wrote %s (%zu bytes)\n", argv[2], lay.size());
	return 0;
}
