// tools/ldpc_layout_opt.cpp -- offline optimiser of the decoder's shared-memory layout: mercury_b200/data/ldpc_tables.bin -> ldpc_layout.bin ('MLAY').
//
//   g++ -O2 -std=c++17 -Imercury_b200/csrc tools/ldpc_layout_opt.cpp -o /tmp/ldpc_layout_opt && /tmp/ldpc_layout_opt mercury_b200/data/ldpc_tables.bin mercury_b200/data/ldpc_layout.bin [moves]
//
// mb_ldpc_kernel keeps posterior[1600] and one message per Tanner-graph edge in shared memory, both as float2 (a PAIR of frames), so a
// gather is a 64-bit access: the hardware serves it half-warp by half-warp over 16 eight-byte banks, and inside a half-warp two lanes
// conflict when they name different words of the same bank.  A warp owns a check task (mb_tables.h: mb_ldpc_split -- a group of 32 sorted
// checks is S tasks of 32 / S checks, S lanes per check, lane l = j * (32 / S) + cl holding edge positions j * Dp + k) or a group of 32
// variables, and at step k every lane gathers through its k-th edge: a check-task lane reads posterior[variable], a variable-group lane
// reads message[slot of the edge].  The bank of a posterior is the variable's position in ITS group of 32 modulo 16; the bank of a
// message is the lane of its slot modulo 16, i.e. (j * (32 / S) + cl) & 15 of the check that owns it.
// Two things are free without touching the kernel or the graph: the ORDER of a node's edges, and the order of equal-degree nodes INSIDE a
// group of 32 -- here restricted to swaps that keep a node in its half-warp (variables, unsplit checks) or in its task (split checks),
// so that the two sides decouple: check-side gathers depend on the checks' edge orders and the variables' in-half positions,
// variable-side gathers on the variables' edge orders and the checks' in-task positions (given the final check-side edge orders).
// This tool minimises the sum over (task / group, step) of the wavefronts by hill climbing with plateau moves (fixed seed) and writes the
// layout file: new node orders and edge orders.  The decoder's arithmetic is unchanged up to the order in which fp32 sums run;
// mb_tables.cpp validates the file (permutations of the reference rows) and uses the reference order without it.  The tables file itself
// (what tools/extract_ldpc_tables.py derives from the reference) is not touched.
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <string>
#include <vector>

#include "mb_tables.h"

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

struct Graph {
	int N, P;
	std::vector<std::vector<int>> crow, vrow;
	std::vector<int> cat, vat;    // node at sorted position i
	std::vector<int> cpos, vpos;  // sorted position of node
	std::vector<int> gS, gDp;     // per check group: lanes per check, edges per lane
	int S(int c) const { return gS[cpos[c] >> 5]; }
	int Dp(int c) const { return gDp[cpos[c] >> 5]; }
	int per(int c) const { return 32 / S(c); }
	int task(int c) const { return (cpos[c] & 31) / per(c); }
	int cl(int c) const { return (cpos[c] & 31) % per(c); }
};

// wavefronts of one 64-bit gather: per half-warp the largest number of distinct words in one 8-byte bank
struct SetCost {
	int cnt[2][16] = {{0}}, m[2] = {0, 0};
	int seen[2][32], ns[2] = {0, 0};
	void add(int half, int bank, int word)  // word < 0: never equal to another lane's
	{
		if (word >= 0) {
			for (int i = 0; i < ns[half]; i++)
				if (seen[half][i] == word) return;  // the same word read twice is a broadcast, not a conflict
			seen[half][ns[half]++] = word;
		}
		m[half] = std::max(m[half], ++cnt[half][bank]);
	}
	int wavefronts() const { return m[0] + m[1]; }
	int objective() const  // 64 * wavefronts (what the hardware pays) + squared bank loads (a slope across the plateaus)
	{
		int sq = 0;
		for (int h = 0; h < 2; h++)
			for (int b = 0; b < 16; b++) sq += cnt[h][b] * cnt[h][b];
		return 64 * wavefronts() + sq;
	}
};

// ---- check side: set (g, t, k) --------------------------------------------------------------------------------------------------------
struct CheckSide {
	Graph &G;
	std::vector<int> setbase;  // first set id of each check group
	std::vector<int> cost;
	explicit CheckSide(Graph &g) : G(g)
	{
		int n = 0;
		for (size_t g2 = 0; g2 < G.gS.size(); g2++) setbase.push_back(n), n += G.gS[g2] * G.gDp[g2];
		cost.assign(n, 0);
		for (int i = 0; i < n; i++) cost[i] = eval(i).objective();
	}
	int set_of(int c, int p) const { return setbase[G.cpos[c] >> 5] + G.task(c) * G.Dp(c) + p % G.Dp(c); }
	SetCost eval(int sid) const
	{
		int g = (int)(std::upper_bound(setbase.begin(), setbase.end(), sid) - setbase.begin()) - 1;
		const int S = G.gS[g], Dp = G.gDp[g], per = 32 / S, t = (sid - setbase[g]) / Dp, k = (sid - setbase[g]) % Dp;
		SetCost sc;
		for (int cl = 0; cl < per; cl++) {
			const int pos = g * 32 + t * per + cl;
			for (int j = 0; j < S; j++) {
				const int half = (j * per + cl) / 16, p = j * Dp + k;
				if (pos >= G.P || p >= (int)G.crow[G.cat[pos]].size()) sc.add(half, 0, 1 << 20);  // padding: the one neutral word, bank of position 0
				else {
					const int v = G.crow[G.cat[pos]][p];
					sc.add(half, G.vpos[v] & 15, v);
				}
			}
		}
		return sc;
	}
	long wavefronts() const
	{
		long s = 0;
		for (size_t i = 0; i < cost.size(); i++) s += eval((int)i).wavefronts();
		return s;
	}
	void run(long moves)
	{
		std::vector<int> nodes;
		for (int c = 0; c < G.P; c++)
			if (G.crow[c].size() >= 2) nodes.push_back(c);
		for (long it = 0; it < moves; it++) {
			if (rnd() & 1) {  // swap two edges of one check
				const int c = nodes[rnd() % nodes.size()];
				std::vector<int> &r = G.crow[c];
				const int i = rnd() % r.size();
				int j = rnd() % (r.size() - 1);
				if (j >= i) j++;
				const int si = set_of(c, i), sj = set_of(c, j);
				const int old = cost[si] + (sj != si ? cost[sj] : 0);
				std::swap(r[i], r[j]);
				const int ci = eval(si).objective(), cj = sj != si ? eval(sj).objective() : 0;
				if (ci + cj <= old) {
					cost[si] = ci;
					if (sj != si) cost[sj] = cj;
				} else
					std::swap(r[i], r[j]);
			} else {  // swap two variables of equal degree inside one half of one variable group
				const int a = rnd() % G.N;
				const int base = G.vpos[a] & ~15;
				const int pb = base + (int)(rnd() % 16);
				if (pb >= G.N) continue;
				const int b = G.vat[pb];
				if (a == b || G.vrow[a].size() != G.vrow[b].size()) continue;
				int aff[64], na = 0;
				for (int v : {a, b})
					for (int c : G.vrow[v]) {
						const std::vector<int> &r = G.crow[c];
						const int sid = set_of(c, (int)(std::find(r.begin(), r.end(), v) - r.begin()));
						bool dup = false;
						for (int q = 0; q < na; q++) dup |= aff[q] == sid;
						if (!dup && na < 64) aff[na++] = sid;
					}
				int old = 0, neu = 0, nc[64];
				for (int q = 0; q < na; q++) old += cost[aff[q]];
				std::swap(G.vpos[a], G.vpos[b]), G.vat[G.vpos[a]] = a, G.vat[G.vpos[b]] = b;
				for (int q = 0; q < na; q++) neu += nc[q] = eval(aff[q]).objective();
				if (neu <= old)
					for (int q = 0; q < na; q++) cost[aff[q]] = nc[q];
				else
					std::swap(G.vpos[a], G.vpos[b]), G.vat[G.vpos[a]] = a, G.vat[G.vpos[b]] = b;
			}
		}
	}
};

// ---- variable side: set (vg, k) -------------------------------------------------------------------------------------------------------
struct VarSide {
	Graph &G;
	std::vector<int> setbase, vgdeg;
	std::vector<int> cost;
	std::vector<std::vector<int>> ejl;  // per variable, per edge: j * per of its slot (fixed: the check-side edge orders are final)
	explicit VarSide(Graph &g) : G(g)
	{
		int n = 0;
		for (int vg = 0; vg * 32 < G.N; vg++) {
			const int d = (int)((G.vrow[G.vat[vg * 32]].size() + 1) & ~(size_t)1);
			setbase.push_back(n), vgdeg.push_back(d), n += d;
		}
		ejl.resize(G.N);
		for (int v = 0; v < G.N; v++)
			for (int c : G.vrow[v]) {
				const std::vector<int> &r = G.crow[c];
				const int p = (int)(std::find(r.begin(), r.end(), v) - r.begin());
				ejl[v].push_back(p / G.Dp(c) * G.per(c));
			}
		cost.assign(n, 0);
		for (int i = 0; i < n; i++) cost[i] = eval(i).objective();
	}
	int set_of(int v, int k) const { return setbase[G.vpos[v] >> 5] + k; }
	SetCost eval(int sid) const
	{
		const int vg = (int)(std::upper_bound(setbase.begin(), setbase.end(), sid) - setbase.begin()) - 1, k = sid - setbase[vg];
		SetCost sc;
		for (int i = 0; i < 32 && vg * 32 + i < G.N; i++) {
			const int v = G.vat[vg * 32 + i];
			if (k >= (int)G.vrow[v].size()) sc.add(i / 16, 0, 1 << 20);  // padding: the always-zero message behind the last slot
			else sc.add(i / 16, (ejl[v][k] + G.cl(G.vrow[v][k])) & 15, -1);
		}
		return sc;
	}
	long wavefronts() const
	{
		long s = 0;
		for (size_t i = 0; i < cost.size(); i++) s += eval((int)i).wavefronts();
		return s;
	}
	void run(long moves)
	{
		std::vector<int> nodes;
		for (int v = 0; v < G.N; v++)
			if (G.vrow[v].size() >= 2) nodes.push_back(v);
		for (long it = 0; it < moves; it++) {
			if (rnd() & 1) {  // swap two edges of one variable
				const int v = nodes[rnd() % nodes.size()];
				std::vector<int> &r = G.vrow[v];
				const int i = rnd() % r.size();
				int j = rnd() % (r.size() - 1);
				if (j >= i) j++;
				const int si = set_of(v, i), sj = set_of(v, j);
				const int old = cost[si] + cost[sj];
				std::swap(r[i], r[j]), std::swap(ejl[v][i], ejl[v][j]);
				const int ci = eval(si).objective(), cj = eval(sj).objective();
				if (ci + cj <= old) cost[si] = ci, cost[sj] = cj;
				else std::swap(r[i], r[j]), std::swap(ejl[v][i], ejl[v][j]);
			} else {  // swap two checks of equal degree inside one task (split groups) or one half (unsplit groups)
				const int a = rnd() % G.P;
				const int span = std::min(16, G.per(a));
				const int base = G.cpos[a] - (G.cpos[a] & 31) % span;  // per is 16, 8 or 4 for split groups: tasks are aligned to their size
				const int pb = base + (int)(rnd() % span);
				if (pb >= G.P) continue;
				const int b = G.cat[pb];
				if (a == b || G.crow[a].size() != G.crow[b].size()) continue;
				int aff[128], na = 0;
				for (int c : {a, b})
					for (int v : G.crow[c]) {
						const std::vector<int> &r = G.vrow[v];
						const int sid = set_of(v, (int)(std::find(r.begin(), r.end(), c) - r.begin()));
						bool dup = false;
						for (int q = 0; q < na; q++) dup |= aff[q] == sid;
						if (!dup && na < 128) aff[na++] = sid;
					}
				int old = 0, neu = 0, nc[128];
				for (int q = 0; q < na; q++) old += cost[aff[q]];
				std::swap(G.cpos[a], G.cpos[b]), G.cat[G.cpos[a]] = a, G.cat[G.cpos[b]] = b;
				for (int q = 0; q < na; q++) neu += nc[q] = eval(aff[q]).objective();
				if (neu <= old)
					for (int q = 0; q < na; q++) cost[aff[q]] = nc[q];
				else
					std::swap(G.cpos[a], G.cpos[b]), G.cat[G.cpos[a]] = a, G.cat[G.cpos[b]] = b;
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
		Graph G;
		G.N = r.h[1], G.P = r.h[3];
		G.crow = r.crow, G.vrow = r.vrow;
		G.cat.resize(G.P), G.vat.resize(G.N), G.cpos.resize(G.P), G.vpos.resize(G.N);
		std::iota(G.cat.begin(), G.cat.end(), 0);
		std::stable_sort(G.cat.begin(), G.cat.end(), [&](int a, int b) { return G.crow[a].size() > G.crow[b].size(); });
		std::iota(G.vat.begin(), G.vat.end(), 0);
		std::stable_sort(G.vat.begin(), G.vat.end(), [&](int a, int b) { return G.vrow[a].size() > G.vrow[b].size(); });
		for (int i = 0; i < G.P; i++) G.cpos[G.cat[i]] = i;
		for (int i = 0; i < G.N; i++) G.vpos[G.vat[i]] = i;
		for (int g = 0; g * 32 < G.P; g++) {
			int S, Dp;
			mb_ldpc_split((int)G.crow[G.cat[g * 32]].size(), &S, &Dp);
			G.gS.push_back(S), G.gDp.push_back(Dp);
		}
		CheckSide cs(G);
		const long c0 = cs.wavefronts();
		cs.run(moves);
		const long c1 = cs.wavefronts();
		VarSide vs(G);  // after the check side: the lanes of the slots depend on the final check-side edge orders
		const long v0 = vs.wavefronts();
		vs.run(moves);
		const long v1 = vs.wavefronts();
		printf("rate %2d/16: check-side gathers %ld -> %ld wavefronts over %zu steps (%.2f -> %.2f), variable-side %ld -> %ld over %zu (%.2f -> %.2f); minimum 2\n",
		       r.h[0], c0, c1, cs.cost.size(), (double)c0 / cs.cost.size(), (double)c1 / cs.cost.size(), v0, v1, vs.cost.size(), (double)v0 / vs.cost.size(),
		       (double)v1 / vs.cost.size());
		put16(r.h[0]), put16((uint16_t)G.N), put16((uint16_t)G.P), put16(0), put32(r.ne);
		for (int v : G.vat) put16((uint16_t)v);
		for (int c : G.cat) put16((uint16_t)c);
		for (int c = 0; c < G.P; c++)
			for (int v : G.crow[c]) put16((uint16_t)v);
		for (int v = 0; v < G.N; v++)
			for (int c : G.vrow[v]) put16((uint16_t)c);
	}
	f = fopen(argv[2], "wb");
	if (!f) return perror(argv[2]), 1;
	fwrite(lay.data(), 1, lay.size(), f);
	fclose(f);
	printf("wrote %s (%zu bytes)\n", argv[2], lay.size());
	return 0;
}
