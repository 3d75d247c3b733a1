// mb_tables.cpp -- host-side construction of the table blob (see mb_tables.h).
//
// Init-time work only ("deterministic tables, per mode"): nothing here runs per frame.
// The tables are re-derived from first principles (PRNG, lattice rule, interleaver index algebra) rather than
// dumped from the reference, and cross-checked against the oracle in tests/test_tables.py.
#include "mb_tables.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <numeric>

namespace {

// ---- glibc random() TYPE_3 (reference: source/common/os_interop.cc:192-283) -------------------------------
// r[i] = r[i-3] + r[i-31] mod 2^32, Park-Miller seeding, 310 warm-up draws, output >> 1; seed 0 -> 1.
}  // namespace

void mb_srandom(uint32_t st[35], unsigned seed)
{
	if (seed == 0) seed = 1;
	int32_t r = (int32_t)seed;
	st[0] = (uint32_t)r;
	for (int i = 1; i < 31; i++) {
		int64_t w = 16807LL * (r % 127773) - 2836LL * (r / 127773);
		if (w < 0) w += 2147483647;
		r = (int32_t)w;
		st[i] = (uint32_t)r;
	}
	for (int i = 31; i < 34; i++) st[i] = st[i - 31];
	st[34] = 0;  // ring cursor
	for (int i = 0; i < 310; i++) (void)mb_random(st);
}

int mb_random(uint32_t st[35])
{
	uint32_t i = st[34];
	uint32_t v = st[(i + 3) % 34] + st[(i + 31) % 34];
	st[i] = v;
	st[34] = (i + 1) % 34;
	return (int)(v >> 1);
}

int mb_rate_index(int rate_num)
{
	static const int rates[MB_NRATES] = {1, 2, 3, 4, 5, 6, 8, 14};
	for (int i = 0; i < MB_NRATES; i++)
		if (rates[i] == rate_num) return i;
	return -1;
}

namespace {

struct ModeDef {
	int M, rate, pre, est;
};
// CONFIG_0..16: telecom_system.cc:2506-2624
const ModeDef kModes[MB_NMODES] = {
	{2, 1, 4, 1},  {2, 2, 4, 1},  {2, 3, 4, 1},  {2, 4, 4, 1},  {2, 5, 4, 1},  {2, 6, 4, 1},
	{2, 8, 4, 1},  {4, 5, 4, 1},  {4, 6, 4, 1},  {4, 8, 4, 1},  {8, 6, 3, 1},  {8, 8, 3, 1},
	{4, 14, 3, 1}, {16, 8, 2, 1}, {8, 14, 2, 1}, {16, 14, 2, 0}, {32, 14, 1, 0},
};

int nsymb_of(int M)  // telecom_system.cc:1818-1826 (HIGH_DENSITY pilots, the compiled default)
{
	switch (M) {
	case 2: return 48;
	case 4: return 24;
	case 8: return 16;
	case 16: return 12;
	case 32: return 9;
	}
	return 0;
}

struct Blob {
	std::vector<uint8_t> b;
	uint32_t reserve(size_t bytes, size_t align = 16)
	{
		size_t off = (b.size() + align - 1) / align * align;
		b.resize(off + bytes, 0);
		return (uint32_t)off;
	}
	template <typename T>
	uint32_t put(const std::vector<T> &v)
	{
		uint32_t off = reserve(v.size() * sizeof(T));
		if (!v.empty()) memcpy(b.data() + off, v.data(), v.size() * sizeof(T));
		return off;
	}
};

struct RateTables {
	int rate_num = 0, N = 0, K = 0, P = 0, n_edges = 0;
	std::vector<std::vector<int>> crow;  // check -> variables, reference order (ascending) unless ldpc_layout.bin re-orders it
	std::vector<std::vector<int>> vrow;  // variable -> checks, reference V-row order unless ldpc_layout.bin re-orders it
	std::vector<int> corder, vorder;     // optional ('MLAY' section): checks / variables in decoder order (degree descending)
};

// Optional layout file mercury_b200/data/ldpc_layout.bin ('MLAY', written by tools/ldpc_layout_opt.cpp): for every rate the order of the nodes inside their
// warp groups and the order of every node's edges, chosen offline so that the decoder's shared-memory gathers hit distinct banks.
// It only PERMUTES what the reference tables define: every row must be a permutation of the reference row and the node orders
// permutations that keep the degrees descending, else the section is rejected as a whole.
bool apply_layout_section(FILE *f, std::vector<RateTables> &rates, std::string &err)
{
	uint8_t hdr[12];
	const size_t got = fread(hdr, 1, 12, f);
	if (got == 0) return true;  // empty file: reference order
	uint32_t ver, n;
	memcpy(&ver, hdr + 4, 4), memcpy(&n, hdr + 8, 4);
	if (got != 12 || memcmp(hdr, "MLAY", 4) != 0 || ver != 1 || n != rates.size()) {
		err = "bad LDPC layout section";
		return false;
	}
	for (RateTables &t : rates) {
		uint16_t h[4];
		uint32_t ne;
		if (fread(h, 2, 4, f) != 4 || fread(&ne, 4, 1, f) != 1 || h[0] != t.rate_num || h[1] != t.N || h[2] != t.P || (int)ne != t.n_edges) {
			err = "LDPC layout section does not match the tables";
			return false;
		}
		std::vector<uint16_t> vo(t.N), co(t.P), ce(ne), ve(ne);
		if (fread(vo.data(), 2, t.N, f) != (size_t)t.N || fread(co.data(), 2, t.P, f) != (size_t)t.P || fread(ce.data(), 2, ne, f) != ne ||
		    fread(ve.data(), 2, ne, f) != ne) {
			err = "LDPC layout section truncated";
			return false;
		}
		auto order_ok = [&](const std::vector<uint16_t> &o, const std::vector<std::vector<int>> &rows) {
			std::vector<char> seen(rows.size(), 0);
			for (size_t i = 0; i < o.size(); i++) {
				if (o[i] >= rows.size() || seen[o[i]]) return false;
				seen[o[i]] = 1;
				if (i > 0 && rows[o[i]].size() > rows[o[i - 1]].size()) return false;
			}
			return true;
		};
		auto rows_ok = [&](const std::vector<uint16_t> &flat, std::vector<std::vector<int>> &rows) {
			size_t e = 0;
			std::vector<std::vector<int>> neu(rows.size());
			for (size_t r = 0; r < rows.size(); r++) {
				neu[r].assign(flat.begin() + (long)e, flat.begin() + (long)(e + rows[r].size()));
				e += rows[r].size();
				std::vector<int> a = neu[r], b = rows[r];
				std::sort(a.begin(), a.end()), std::sort(b.begin(), b.end());
				if (a != b) return false;
			}
			rows.swap(neu);
			return true;
		};
		if (!order_ok(vo, t.vrow) || !order_ok(co, t.crow) || !rows_ok(ce, t.crow) || !rows_ok(ve, t.vrow)) {
			err = "LDPC layout section is not a permutation of the reference tables";
			return false;
		}
		t.vorder.assign(vo.begin(), vo.end()), t.corder.assign(co.begin(), co.end());
	}
	return true;
}

bool load_ldpc_file(const char *path, std::vector<RateTables> &rates, std::string &err)
{
	FILE *f = fopen(path, "rb");
	if (!f) {
		err = std::string("cannot open LDPC table file ") + path;
		return false;
	}
	uint8_t hdr[12];
	if (fread(hdr, 1, 12, f) != 12 || memcmp(hdr, "MLDP", 4) != 0) {
		fclose(f);
		err = "bad LDPC table file header";
		return false;
	}
	uint32_t n;
	memcpy(&n, hdr + 8, 4);
	for (uint32_t r = 0; r < n; r++) {
		uint16_t h[6];
		uint32_t ne;
		if (fread(h, 2, 6, f) != 6 || fread(&ne, 4, 1, f) != 1) break;
		RateTables t;
		t.rate_num = h[0], t.N = h[1], t.K = h[2], t.P = h[3], t.n_edges = (int)ne;
		std::vector<uint16_t> cdeg(t.P), ev(ne), vdeg(t.N), vc(ne);
		if (fread(cdeg.data(), 2, t.P, f) != (size_t)t.P || fread(ev.data(), 2, ne, f) != ne ||
		    fread(vdeg.data(), 2, t.N, f) != (size_t)t.N || fread(vc.data(), 2, ne, f) != ne)
			break;
		size_t e = 0;
		t.crow.resize(t.P);
		for (int c = 0; c < t.P; c++)
			for (int j = 0; j < cdeg[c]; j++) t.crow[c].push_back(ev[e++]);
		e = 0;
		t.vrow.resize(t.N);
		for (int v = 0; v < t.N; v++)
			for (int j = 0; j < vdeg[v]; j++) t.vrow[v].push_back(vc[e++]);
		rates.push_back(std::move(t));
	}
	if (rates.size() != MB_NRATES) {
		fclose(f);
		err = "LDPC table file truncated";
		return false;
	}
	fclose(f);
	// optional layout file next to the tables: <dir>/ldpc_layout.bin
	std::string lp(path);
	const size_t slash = lp.find_last_of('/');
	lp = (slash == std::string::npos ? std::string() : lp.substr(0, slash + 1)) + "ldpc_layout.bin";
	FILE *lf = fopen(lp.c_str(), "rb");
	if (!lf) return true;  // none: reference order
	const bool ok = apply_layout_section(lf, rates, err);
	fclose(lf);
	return ok;
}

std::string build_rate(Blob &bl, const RateTables &t, MbRate &out, std::vector<uint16_t> &var_of_cw)
{
	const int N = t.N, P = t.P;
	if (N != MB_N) return "unexpected LDPC block length";
	// checks sorted by degree, descending (stable)
	std::vector<int> csorted(P);
	std::iota(csorted.begin(), csorted.end(), 0);
	std::stable_sort(csorted.begin(), csorted.end(), [&](int a, int b) { return t.crow[a].size() > t.crow[b].size(); });
	if (!t.corder.empty()) csorted = t.corder;  // layout section: same degree order, bank-friendly positions inside the warp groups
	std::vector<int> cpos(P);
	for (int i = 0; i < P; i++) cpos[csorted[i]] = i;
	// variables renumbered by degree, descending (stable)
	std::vector<int> vsorted(N);
	std::iota(vsorted.begin(), vsorted.end(), 0);
	std::stable_sort(vsorted.begin(), vsorted.end(), [&](int a, int b) { return t.vrow[a].size() > t.vrow[b].size(); });
	if (!t.vorder.empty()) vsorted = t.vorder;
	var_of_cw.assign(N, 0);
	for (int i = 0; i < N; i++) var_of_cw[vsorted[i]] = (uint16_t)i;

	int max_cdeg = (int)t.crow[csorted[0]].size(), max_vdeg = (int)t.vrow[vsorted[0]].size();
	if (max_cdeg > MB_MAX_CDEG || max_vdeg > MB_MAX_VDEG) return "LDPC degree exceeds compiled limits";
	// warp-blocked ELL: groups of 32 sorted checks / variables, each padded to the degree of its first member
	std::vector<uint32_t> cgbase(MB_MAX_GROUPS, 0), vgbase(MB_MAX_GROUPS, 0);
	uint32_t c_slots = 0, v_slots = 0;
	for (int g = 0; g * 32 < P; g++) {  // S tasks of Dp steps each (mb_ldpc_split): 32 * S * Dp >= 32 * d slots
		int S, Dp;
		mb_ldpc_split((int)t.crow[csorted[g * 32]].size(), &S, &Dp);
		cgbase[g] = c_slots;
		c_slots += 32u * (uint32_t)(S * Dp);
	}
	for (int g = 0; g * 32 < N; g++) {  // variable groups are padded to an EVEN degree: the kernel's loop takes two edges per trip
		vgbase[g] = v_slots;
		v_slots += 32u * (uint32_t)((t.vrow[vsorted[g * 32]].size() + 1) & ~(size_t)1);
	}
	if (c_slots >= 0xFFFFu || v_slots >= 0xFFFFu) return "LDPC slot count exceeds 16-bit slot ids";
	auto cslot = [&](int k, int cs) { return mb_ldpc_cslot(cgbase.data(), (int)t.crow[csorted[cs & ~31]].size(), cs, k); };
	auto vslot = [&](int k, int vs) { return vgbase[vs >> 5] + 32u * (uint32_t)k + (uint32_t)(vs & 31); };

	std::vector<uint8_t> cdeg(P), vdeg(N);
	std::vector<uint16_t> edge_var(c_slots, 0xFFFF), vedge(v_slots, (uint16_t)c_slots), check_of_sorted(P);
	// variable-side padding points at slot c_slots: one extra always-zero message, so a whole group can run to its largest degree
	int placed = 0;
	for (int i = 0; i < P; i++) {
		int c = csorted[i];
		cdeg[i] = (uint8_t)t.crow[c].size();
		check_of_sorted[i] = (uint16_t)c;
		for (size_t k = 0; k < t.crow[c].size(); k++) edge_var[cslot((int)k, i)] = var_of_cw[t.crow[c][k]], placed++;
	}
	if (placed != t.n_edges) return "LDPC edge count mismatch";
	placed = 0;
	for (int i = 0; i < N; i++) {
		int v = vsorted[i];
		vdeg[i] = (uint8_t)t.vrow[v].size();
		for (size_t k = 0; k < t.vrow[v].size(); k++) {
			int c = t.vrow[v][k];
			const std::vector<int> &row = t.crow[c];
			auto it = std::find(row.begin(), row.end(), v);
			if (it == row.end()) return "LDPC V/C tables inconsistent";
			int kc = (int)(it - row.begin());
			vedge[vslot((int)k, i)] = (uint16_t)cslot(kc, cpos[c]);
			placed++;
		}
	}
	if (placed != t.n_edges) return "LDPC edge count mismatch (variable side)";
	out.rate_num = t.rate_num, out.N = N, out.K = t.K, out.P = P, out.n_edges = t.n_edges;
	out.max_cdeg = max_cdeg, out.max_vdeg = max_vdeg, out.c_slots = (int32_t)c_slots, out.v_slots = (int32_t)v_slots;
	out.reserved = 0;
	out.off_cdeg = bl.put(cdeg);
	out.off_cgbase = bl.put(cgbase);
	out.off_edge_var = bl.put(edge_var);
	out.off_vdeg = bl.put(vdeg);
	out.off_vgbase = bl.put(vgbase);
	{
		std::vector<uint8_t> vgdeg(MB_MAX_GROUPS, 0);
		for (int g = 0; g * 32 < N; g++) vgdeg[g] = (uint8_t)((t.vrow[vsorted[g * 32]].size() + 1) & ~(size_t)1);
		out.off_vgdeg = bl.put(vgdeg);
	}
	out.off_vedge = bl.put(vedge);
	{
		std::vector<uint16_t> evb(c_slots, 0), veb(v_slots, 0);
		for (uint32_t i = 0; i < c_slots; i++) evb[i] = (uint16_t)((edge_var[i] == 0xFFFF ? (uint32_t)N : (uint32_t)edge_var[i]) * 8);
		for (uint32_t i = 0; i < v_slots; i++) veb[i] = (uint16_t)(vedge[i] * 8);
		if (c_slots * 8 > 0xFFFFu) return "LDPC slot byte offsets exceed 16 bits";
		out.off_edge_varb = bl.put(evb);
		out.off_vedgeb = bl.put(veb);
		// tail of degree-<=2 variables (sorted by descending degree, so it is a suffix), from a multiple of 32
		int tail = N;
		while (tail > 0 && t.vrow[vsorted[tail - 1]].size() <= 2) tail--;
		tail = (tail + 31) / 32 * 32;
		out.vtail_start = tail;
		std::vector<uint32_t> vtail((size_t)(N - tail), 0);
		for (int i = tail; i < N; i++) {
			uint32_t o[2] = {c_slots * 8u, c_slots * 8u};
			for (size_t k = 0; k < t.vrow[vsorted[i]].size(); k++) o[k] = (uint32_t)vedge[vslot((int)k, i)] * 8u;
			vtail[(size_t)(i - tail)] = o[0] | (o[1] << 16);
		}
		out.off_vtail = bl.put(vtail);
		// longest-processing-time-first assignment of tasks to warps; cost = what balances the warps: warp instructions of the task in
		// the kernel (mb_ldpc.cu: ~25 per edge of a pair of frames in an unrolled body, ~12 more per edge and ~40 per task when the
		// check is split over lanes, a degree-2 check is a pass-through; ~3 per edge on the variable side)
		// check_side: each warp's tasks are then sorted by body (edges per lane + 1 for a split task), and the last word of the warp's row
		// holds the number of tasks per body, 4 bits each from body 2: the kernel runs one loop per body instead of a switch per task.
		auto lpt = [&](std::vector<std::pair<int, uint32_t>> items, uint32_t &off, bool check_side) -> bool {  // (cost, descriptor)
			std::stable_sort(items.begin(), items.end(), [](const std::pair<int, uint32_t> &a, const std::pair<int, uint32_t> &b) { return a.first > b.first; });
			std::vector<uint32_t> sched(MB_LDPC_WARPS * MB_SCHED_LEN, 0u);
			int load[MB_LDPC_WARPS] = {0}, cnt[MB_LDPC_WARPS] = {0};
			for (const auto &it : items) {
				int w = 0;
				for (int i = 1; i < MB_LDPC_WARPS; i++)
					if (load[i] < load[w]) w = i;
				if (cnt[w] >= MB_SCHED_LEN - 2) return false;
				sched[w * MB_SCHED_LEN + cnt[w]++] = it.second;
				load[w] += it.first;
			}
			if (check_side)
				for (int w = 0; w < MB_LDPC_WARPS; w++) {
					uint32_t *row = sched.data() + w * MB_SCHED_LEN;
					std::stable_sort(row, row + cnt[w], [](uint32_t a, uint32_t b) { return MB_CDESC_BODY(a) < MB_CDESC_BODY(b); });
					uint32_t counts = 0;
					for (int i = 0; i < cnt[w]; i++) {
						const uint32_t body = MB_CDESC_BODY(row[i]);
						if (body < 2 || body > MB_LDPC_DMAX + 1 || ((counts >> (4 * (body - 2))) & 15u) == 15u) return false;
						counts += 1u << (4 * (body - 2));
					}
					row[MB_SCHED_LEN - 1] = counts;
				}
			off = bl.put(sched);
			return true;
		};
		int kc[5] = {26, 25, 15, 37, 55};
		if (const char *e = getenv("MERCURY_B200_LDPC_COST")) sscanf(e, "%d,%d,%d,%d,%d", &kc[0], &kc[1], &kc[2], &kc[3], &kc[4]);  // tuning only
		std::vector<std::pair<int, uint32_t>> ctasks, vtasks;
		for (int g = 0; g * 32 < P; g++) {
			int S, Dp, l2 = 0;
			mb_ldpc_split((int)t.crow[csorted[g * 32]].size(), &S, &Dp);
			while ((1 << l2) < S) l2++;
			for (int tk = 0; tk < S; tk++) {
				if (g * 32 + tk * (32 / S) >= P) break;  // no check left for this task (last group)
				const uint32_t base = cgbase[g] + (uint32_t)(tk * Dp * 32);
				if (base > 0xFFFFu || Dp > 15 || g + 1 > 127) return "check schedule overflow";
				const int cost = Dp <= 2 && S == 1 ? kc[0] : (S == 1 ? kc[1] * Dp + kc[2] : kc[3] * Dp + kc[4]);
				ctasks.emplace_back(cost, base | ((uint32_t)Dp << 16) | ((uint32_t)l2 << 20) | ((uint32_t)tk << 22) | ((uint32_t)(g + 1) << 25));
			}
		}
		for (int g = 0; g < tail / 32; g++) {
			const int d = (int)((t.vrow[vsorted[g * 32]].size() + 1) & ~(size_t)1);
			if (vgbase[g] > 0xFFFFu || d > 255) return "variable schedule overflow";
			vtasks.emplace_back(3 * d + 8, vgbase[g] | ((uint32_t)d << 16) | ((uint32_t)(g + 1) << 24));
		}
		if (!lpt(ctasks, out.off_csched, true)) return "check schedule overflow";
		if (!lpt(vtasks, out.off_vsched, false)) return "variable schedule overflow";
	}
	out.off_var_of_cw = bl.put(var_of_cw);
	out.off_check_of_sorted = bl.put(check_of_sorted);
	return "";
}

// constellations: psk.cc:65-226; unit mean power with a *float* normaliser: psk.cc:229-256
void build_constellation(int M, std::vector<float> &out)
{
	static const signed char q16[16][2] = {{-3, 3}, {-3, 1}, {-3, -3}, {-3, -1}, {-1, 3}, {-1, 1}, {-1, -3}, {-1, -1},
					       {3, 3},	{3, 1},	 {3, -3},  {3, -1},  {1, 3},  {1, 1},  {1, -3},	 {1, -1}};
	static const signed char q32[32][2] = {{-3, 5}, {-1, 5}, {-3, -5}, {-1, -5}, {-5, 3}, {-5, 1}, {-5, -3}, {-5, -1},
					       {-1, 3}, {-1, 1}, {-1, -3}, {-1, -1}, {-3, 3}, {-3, 1}, {-3, -3}, {-3, -1},
					       {3, 5},	{1, 5},	 {3, -5},  {1, -5},  {5, 3},  {5, 1},  {5, -3},	 {5, -1},
					       {1, 3},	{1, 1},	 {1, -3},  {1, -1},  {3, 3},  {3, 1},  {3, -3},	 {3, -1}};
	std::vector<double> re(M), im(M);
	const double h = std::sqrt(2.0) / 2.0;
	if (M == 2) {
		re = {1, -1}, im = {0, 0};
	} else if (M == 4) {
		re = {-1, -1, 1, 1}, im = {1, -1, 1, -1};
	} else if (M == 8) {
		re = {-h, -1, 0, -h, 0, h, h, 1}, im = {-h, 0, 1, h, -1, -h, h, 0};
	} else if (M == 16) {
		for (int i = 0; i < 16; i++) re[i] = q16[i][0], im[i] = q16[i][1];
	} else {
		for (int i = 0; i < 32; i++) re[i] = q32[i][0], im[i] = q32[i][1];
	}
	float pn = 0;
	for (int i = 0; i < M; i++) pn = (float)((double)pn + re[i] * re[i] + im[i] * im[i]);
	pn = (float)(1.0 / std::sqrt((double)(pn / M)));
	out.resize(2 * M);
	for (int i = 0; i < M; i++) {
		out[2 * i] = (float)(re[i] * (double)pn);
		out[2 * i + 1] = (float)(im[i] * (double)pn);
	}
}

uint16_t crc_zero_byte(uint16_t s)  // advance the reflected CRC-16/MODBUS register over one zero byte
{
	for (int i = 0; i < 8; i++) s = (s & 1) ? (uint16_t)((s >> 1) ^ 0xA001) : (uint16_t)(s >> 1);
	return s;
}
// The CRC is linear over GF(2): with preset 0 it is the XOR, over the set bits of the message, of the register a message with only
// that bit set leaves after all n_bytes bytes.  The decoder's epilogue XORs these per thread (one byte each) and reduces over the CTA.
std::vector<uint16_t> crc_bit_table(int n_bytes)
{
	std::vector<uint16_t> t(8 * (size_t)n_bytes, 0);
	for (int j = 0; j < n_bytes; j++)
		for (int b = 0; b < 8; b++) {
			uint16_t s = (uint16_t)(1u << b);  // register 0 ^ the byte, then its eight shifts = crc_zero_byte of it
			s = crc_zero_byte(s);
			for (int k = j + 1; k < n_bytes; k++) s = crc_zero_byte(s);
			t[8 * (size_t)j + b] = s;
		}
	return t;
}

std::string build_mode(Blob &bl, int cfg, const MbRate &rate, const std::vector<uint16_t> &var_of_cw, MbMode &m)
{
	const ModeDef &d = kModes[cfg];
	memset(&m, 0, sizeof(m));
	m.config = cfg, m.M = d.M, m.rate_num = d.rate, m.rate_idx = mb_rate_index(d.rate);
	m.bps = 0;
	while ((1 << m.bps) < d.M) m.bps++;
	m.Nsymb = nsymb_of(d.M), m.estimator = d.est, m.preamble_nSymb = d.pre;
	m.phase_only = (d.M == 2 || d.M == 4 || d.M == 8);  // telecom_system.cc:2647-2654
	m.boost = 1.33f;				    // physical_config.cc:46
	m.K = rate.K, m.P = rate.P;
	const int S = m.Nsymb, C = MB_NC, cells = S * C;

	// pilot lattice (ofdm.cc:976-1064 with Dx=1, Dy=3, edges DATA): pilot iff s%3 == c%3.  The kernels rely on
	// this closed form (row sums step by 3), so it is built by the reference's walk and then verified.
	std::vector<uint8_t> is_pilot(cells, 0);
	{
		const int ncm = std::max(C, S), Dx = 1, Dy = 3;
		std::vector<uint8_t> v((size_t)ncm * ncm, 0);
		for (int x = 0, y = 0; x < ncm && y < ncm; y++, x += Dx) {
			for (int j = y; j < ncm; j += Dy) v[j * ncm + x] = 1;
			for (int j = y; j >= 0; j -= Dy) v[j * ncm + x] = 1;
		}
		int cnt = 0;
		for (int j = 0; j < S; j++) cnt += v[j * ncm + C - 1];
		if (cnt < 2)
			for (int j = 0; j < ncm; j++) v[j * ncm + C - 1] = v[j * ncm];
		for (int s = 0; s < S; s++)
			for (int c = 0; c < C; c++) {
				is_pilot[s * C + c] = v[s * ncm + c];
				if ((v[s * ncm + c] != 0) != ((s % 3) == (c % 3))) return "pilot lattice is not s%3==c%3";
			}
	}
	std::vector<float> pval(cells, 0.f), pinv(cells, 0.f), invn(cells, 0.f);
	std::vector<uint16_t> pilot_cell, data_cell;
	{
		uint32_t st[35];
		mb_srandom(st, 0);  // ofdm.cc:940-951, seed physical_config.cc:47
		int last = 0;
		for (int i = 0; i < cells; i++) {
			if (!is_pilot[i]) {
				data_cell.push_back((uint16_t)i);
				continue;
			}
			int pv = (mb_random(st) % 2) ^ last;
			last = pv;
			double p = (double)(2 * pv - 1) * (double)m.boost;
			pval[i] = (float)p;
			pinv[i] = (float)(1.0 / p);
			pilot_cell.push_back((uint16_t)i);
		}
	}
	m.nPilots = (int)pilot_cell.size();
	m.nData = (int)data_cell.size();
	for (int s = 0; s < S; s++)  // pilots inside the clipped 21x21 window (ofdm.cc:1364-1390)
		for (int c = 0; c < C; c++) {
			if (!is_pilot[s * C + c]) continue;
			int n = 0;
			for (int k = std::max(0, s - MB_LS_HALF); k <= std::min(S - 1, s + MB_LS_HALF); k++)
				for (int l = std::max(0, c - MB_LS_HALF); l <= std::min(C - 1, c + MB_LS_HALF); l++) n += is_pilot[k * C + l];
			invn[s * C + c] = (float)(1.0 / n);
		}
	m.nBits = m.nData * m.bps;  // data_container.cc:90-172
	m.nReal = m.nBits - m.P;
	m.nVirtual = MB_N - m.nBits;
	m.frame_bytes = (m.nReal - 16) / 8;  // telecom_system.cc:332-335
	if (m.nReal <= 16 || m.nVirtual < 0 || m.nReal + m.nVirtual != m.K) return "inconsistent frame geometry";

	// symbol q of the demapper input <- grid cell: deframer (ofdm.cc:837-852) then complex de-interleaver
	// (interleaver.cc:94-109, block nData/10, telecom_system.cc:2911)
	std::vector<uint16_t> sym_cell(m.nData);
	{
		int bs = m.nData / 10, nb = m.nData / bs;
		for (int q = 0; q < m.nData; q++) {
			int src = q;
			if (q < nb * bs) src = (q % bs) * nb + q / bs;
			sym_cell[q] = data_cell[src];
		}
	}
	// LLR i of the demapper -> float de-interleaver (interleaver.cc:77-92, block nBits/10, :2910) -> expand
	// (telecom_system.cc:1300-1308) -> internal variable numbering of the decoder
	std::vector<uint16_t> dst(m.nBits), dst2(m.nBits, MB_NO_DST);
	{
		int bs = m.nBits / 10, nb = m.nBits / bs;
		for (int i = 0; i < m.nBits; i++) {
			int j = i;
			if (i < nb * bs) j = (i % nb) * bs + i / nb;
			int cw = j < m.nReal ? j : j + m.nVirtual;
			dst[i] = var_of_cw[cw];
			if (j < m.nVirtual) dst2[i] = var_of_cw[m.nReal + j];
		}
	}
	std::vector<float> cons;
	build_constellation(m.M, cons);

	m.crc_bytes = m.nReal / 8;
	std::vector<uint16_t> bit_var(8 * m.crc_bytes);
	std::vector<uint8_t> scr(MB_N);
	{
		uint32_t st[35];
		mb_srandom(st, 0);  // telecom_system.cc:1961-1966 (bit_energy_dispersal_seed = 0), N draws
		for (int i = 0; i < MB_N; i++) scr[i] = (uint8_t)(mb_random(st) % 2);
		for (int i = 0; i < 8 * m.crc_bytes; i++) bit_var[i] = var_of_cw[i];
	}
	// CTA-parallel CRC (crc_bit_table above); the 0xFFFF preset contributes a constant.
	m.crc_reserved = 0;
	std::vector<uint16_t> crcbit = crc_bit_table(m.crc_bytes);
	{
		uint16_t s = 0xFFFF;
		for (int i = 0; i < m.crc_bytes; i++) s = crc_zero_byte(s);
		m.crc_init = s;
	}
	// ---- descriptors of the persistent demodulator kernel: lattice / window / interleaver arithmetic resolved here -------
	// compact pilot rows: row s holds its pilots (columns s%3 + 3j) at [4 + j]; the rest of the 27-wide row is zero, so that
	// every clipped 21-column window is exactly 7 consecutive entries starting at (c + 4 - s%3) / 3.
	m.pinv_mag = std::fabs(pinv[pilot_cell[0]]);
	std::vector<uint32_t> zf_src((size_t)S * MB_ZF_STRIDE, 0);
	for (int s = 0; s < S; s++)
		for (int j = 0; s % 3 + 3 * j < C; j++) {
			const int cell = s * C + s % 3 + 3 * j;
			if (!is_pilot[cell]) return "compact pilot row hits a data cell";
			if (std::fabs(pinv[cell]) != m.pinv_mag) return "pilot magnitudes differ";
			zf_src[(size_t)s * MB_ZF_STRIDE + 4 + j] = (uint32_t)(cell * 8) | (1u << 30) | (pinv[cell] < 0 ? 1u << 31 : 0u);
		}
	std::vector<uint32_t> pilot_rec(4 * (size_t)m.nPilots, 0);
	std::vector<float> pilot_f(2 * (size_t)m.nPilots);
	for (int p = 0; p < m.nPilots; p++) {
		const int cell = pilot_cell[p], s = cell / C, c = cell % C, j = c / 3;
		if (c % 3 != s % 3) return "pilot off the lattice";
		const int zslot = s * MB_ZF_STRIDE + 4 + j;
		const int k0 = std::max(0, s - MB_LS_HALF), k1 = std::min(S - 1, s + MB_LS_HALF);
		for (int r = 0; r < 3; r++) {
			int first = k0 + ((r - k0) % 3 + 3) % 3, last = k1 - ((k1 - r) % 3 + 3) % 3;
			if (first > last || first % 3 != r || last % 3 != r) return "LS window misses a row residue";
			const int col = (c + 4 - r) / 3;  // first compact entry of the clipped 21-column window in a row of residue r
			if (col < 0 || col >= MB_LS_COLS) return "LS window column out of range";
			const uint32_t hi = (uint32_t)(last * MB_LS_COLS + col) * 8u;
			const uint32_t lo = (uint32_t)((first >= 3 ? first - 3 : S) * MB_LS_COLS + (first >= 3 ? col : 0)) * 8u;
			if (hi > 0xFFFFu || lo > 0xFFFFu) return "LS offsets exceed 16 bits";
			pilot_rec[4 * (size_t)p + r] = hi | (lo << 16);
		}
		if (zslot * 8 > 0xFFFF) return "compact slot offset exceeds 16 bits";
		pilot_rec[4 * (size_t)p + 3] = (uint32_t)(cell * 8) | ((uint32_t)(zslot * 8) << 16);
		pilot_f[2 * (size_t)p] = invn[cell];
		pilot_f[2 * (size_t)p + 1] = pval[cell];
	}
	m.data_rec_words = m.bps <= 2 ? 2 : 4;
	std::vector<uint32_t> data_rec((size_t)m.data_rec_words * m.nData, 0);
	{
		std::vector<int> q_of_cell(cells, -1);
		for (int q = 0; q < m.nData; q++) q_of_cell[sym_cell[q]] = q;
		for (int D = 0; D < m.nData; D++) {
			const int cell = data_cell[D], s = cell / C, c = cell % C, q = q_of_cell[cell];
			if (q < 0) return "data cell without a demapper slot";
			const int f = c % 3, last = f + 3 * ((S - 1 - f) / 3);
			const int r0 = s < f ? f : (s > last ? last - 3 : s - ((s - f) % 3));  // interpolator.cc:163-254
			if (r0 < 0 || r0 + 3 > S - 1) return "interpolation rows out of range";
			const int t = s - r0;
			if (t < -2 || t > 5 || !is_pilot[r0 * C + c] || !is_pilot[(r0 + 3) * C + c]) return "interpolation descriptor out of range";
			const uint32_t zs = (uint32_t)(r0 * MB_ZF_STRIDE + 4 + c / 3) * 8u;
			if (cell * 8 >= (1 << 15) || zs >= (1u << 14)) return "data descriptor field overflow";
			uint32_t *rec = &data_rec[(size_t)D * m.data_rec_words];
			rec[0] = (uint32_t)(cell * 8) | (zs << 15) | ((uint32_t)(t + 2) << 29);
			for (int e = 0; e < m.bps; e++) {
				const uint32_t d = MB_HANDOFF((uint32_t)dst[(size_t)q * m.bps + e]) * 4u;
				rec[1 + e / 2] |= d << (16 * (e & 1));
			}
		}
	}
	std::vector<uint16_t> virt;
	for (int i = 0; i < m.nBits; i++)
		if (dst2[i] != MB_NO_DST) {
			virt.push_back((uint16_t)(MB_HANDOFF((uint32_t)dst[i]) * 4));
			virt.push_back((uint16_t)(MB_HANDOFF((uint32_t)dst2[i]) * 4));
		}
	if ((int)virt.size() != 2 * m.nVirtual) return "virtual-bit copy list does not match nVirtual";
	{
		std::vector<uint64_t> neg(S, 0);
		for (int p = 0; p < m.nPilots; p++)
			if (pval[pilot_cell[p]] < 0) neg[pilot_cell[p] / C] |= 1ull << (pilot_cell[p] % C);
		m.off_pilot_neg = bl.put(neg);
	}
	m.off_zf_src = bl.put(zf_src);
	m.off_pilot_rec = bl.put(pilot_rec);
	m.off_pilot_f = bl.put(pilot_f);
	m.off_data_rec = bl.put(data_rec);
	m.off_virt = bl.put(virt);
	m.off_pinv = bl.put(pinv);
	m.off_pval = bl.put(pval);
	m.off_invn = bl.put(invn);
	m.off_pilot_cell = bl.put(pilot_cell);
	m.off_sym_cell = bl.put(sym_cell);
	m.off_llr_dst = bl.put(dst);
	m.off_llr_dst2 = bl.put(dst2);
	m.off_const = bl.put(cons);
	m.off_bit_var = bl.put(bit_var);
	m.off_scr = bl.put(scr);
	m.off_crcbit = bl.put(crcbit);
	return "";
}

}  // namespace

std::string mb_build_blob(const char *ldpc_blob_path, std::vector<uint8_t> &out)
{
	std::vector<RateTables> rt;
	std::string err;
	if (!load_ldpc_file(ldpc_blob_path, rt, err)) return err;
	Blob bl;
	bl.reserve(sizeof(MbBlobHeader));
	MbBlobHeader hdr;
	memset(&hdr, 0, sizeof(hdr));
	hdr.magic = MB_BLOB_MAGIC, hdr.version = MB_BLOB_VERSION;
	{
		// FFT-256 as 16x16: tw[k1*16+n2] = exp(-2 pi i n2 k1 / 256) / 256 (the 1/N of ofdm.cc:439-442 folded in)
		std::vector<float> tw(2 * 256);
		for (int k1 = 0; k1 < 16; k1++)
			for (int n2 = 0; n2 < 16; n2++) {
				double a = -2.0 * M_PI * (double)(n2 * k1) / 256.0;
				tw[2 * (k1 * 16 + n2)] = (float)(std::cos(a) / 256.0);
				tw[2 * (k1 * 16 + n2) + 1] = (float)(std::sin(a) / 256.0);
			}
		hdr.off_twiddle = bl.put(tw);
	}
	std::vector<std::vector<uint16_t>> var_of_cw(MB_NRATES);
	for (int r = 0; r < MB_NRATES; r++) {
		int idx = mb_rate_index(rt[r].rate_num);
		if (idx < 0) return "unknown LDPC rate in table file";
		err = build_rate(bl, rt[r], hdr.rates[idx], var_of_cw[idx]);
		if (!err.empty()) return err;
	}
	for (int c = 0; c < MB_NMODES; c++) {
		int idx = mb_rate_index(kModes[c].rate);
		err = build_mode(bl, c, hdr.rates[idx], var_of_cw[idx], hdr.modes[c]);
		if (!err.empty()) return "mode " + std::to_string(c) + ": " + err;
	}
	bl.reserve(0, 256);
	hdr.total_bytes = (uint32_t)bl.b.size();
	memcpy(bl.b.data(), &hdr, sizeof(hdr));
	out.swap(bl.b);
	return "";
}

std::string mb_validate_blob(const uint8_t *blob, size_t size)
{
	if (size < sizeof(MbBlobHeader)) return "table blob too small";
	MbBlobHeader h;
	memcpy(&h, blob, sizeof(h));
	if (h.magic != MB_BLOB_MAGIC) return "table blob: bad magic";
	if (h.version != MB_BLOB_VERSION) return "table blob: version mismatch";
	if (h.total_bytes != size) return "table blob: size mismatch";
	auto in = [&](uint32_t off, size_t bytes) { return (size_t)off + bytes <= size; };
	if (!in(h.off_twiddle, 2 * 256 * 4)) return "table blob: twiddle out of range";
	for (int r = 0; r < MB_NRATES; r++) {
		const MbRate &t = h.rates[r];
		if (t.N != MB_N || t.n_edges <= 0 || t.n_edges > 8192 || t.P <= 0 || t.P >= MB_N) return "table blob: bad rate record";
		if (t.c_slots < t.n_edges || t.c_slots > 12288 || t.v_slots < t.n_edges || t.v_slots > 12288) return "table blob: bad slot counts";
		if (!in(t.off_cdeg, t.P) || !in(t.off_cgbase, 4 * MB_MAX_GROUPS) || !in(t.off_edge_var, 2 * (size_t)t.c_slots) ||
		    !in(t.off_vdeg, t.N) || !in(t.off_vgbase, 4 * MB_MAX_GROUPS) || !in(t.off_vgdeg, MB_MAX_GROUPS) || !in(t.off_vedge, 2 * (size_t)t.v_slots) ||
		    !in(t.off_var_of_cw, 2 * (size_t)t.N) || !in(t.off_check_of_sorted, 2 * (size_t)t.P) || !in(t.off_edge_varb, 2 * (size_t)t.c_slots) ||
		    !in(t.off_vedgeb, 2 * (size_t)t.v_slots) || !in(t.off_csched, 4 * MB_LDPC_WARPS * MB_SCHED_LEN) || !in(t.off_vsched, 4 * MB_LDPC_WARPS * MB_SCHED_LEN) ||
		    t.vtail_start < 0 || t.vtail_start > t.N || (t.vtail_start & 31) != 0 || !in(t.off_vtail, 4 * (size_t)(t.N - t.vtail_start)))
			return "table blob: rate table out of range";
		// contents the decoder kernel uses as shared-memory byte offsets / loop bounds without further checks
		if ((size_t)t.c_slots * 8 > 0xFFFFu) return "table blob: slot count exceeds the 16-bit byte offsets";
		{
			const uint16_t *evb = reinterpret_cast<const uint16_t *>(blob + t.off_edge_varb), *veb = reinterpret_cast<const uint16_t *>(blob + t.off_vedgeb);
			for (int i = 0; i < t.c_slots; i++)
				if (evb[i] > 8 * MB_N || (evb[i] & 7)) return "table blob: posterior offset out of range";
			for (int i = 0; i < t.v_slots; i++)
				if (veb[i] > 8 * t.c_slots || (veb[i] & 7)) return "table blob: message offset out of range";
			const uint32_t *vt = reinterpret_cast<const uint32_t *>(blob + t.off_vtail);
			for (int i = 0; i < t.N - t.vtail_start; i++)
				if ((vt[i] & 0xFFFFu) > 8u * (uint32_t)t.c_slots || (vt[i] >> 16) > 8u * (uint32_t)t.c_slots || (vt[i] & 0x00070007u)) return "table blob: tail offset out of range";
			const uint32_t *cs = reinterpret_cast<const uint32_t *>(blob + t.off_csched), *vs = reinterpret_cast<const uint32_t *>(blob + t.off_vsched);
			for (int w = 0; w < MB_LDPC_WARPS; w++) {
				if (cs[w * MB_SCHED_LEN + MB_SCHED_LEN - 2] != 0u || vs[w * MB_SCHED_LEN + MB_SCHED_LEN - 1] != 0u) return "table blob: schedule not terminated";
				{  // the per-body counts must describe the row: sorted by body, as many tasks as counted
					uint32_t counts = cs[w * MB_SCHED_LEN + MB_SCHED_LEN - 1], at = 0;
					if (counts >> (4 * MB_LDPC_DMAX)) return "table blob: check schedule counts out of range";
					for (uint32_t body = 2; body <= MB_LDPC_DMAX + 1; body++)
						for (uint32_t n = (counts >> (4 * (body - 2))) & 15u; n > 0; n--, at++)
							if (at >= MB_SCHED_LEN - 2 || cs[w * MB_SCHED_LEN + at] == 0u || MB_CDESC_BODY(cs[w * MB_SCHED_LEN + at]) != body) return "table blob: check schedule counts do not match";
					if (cs[w * MB_SCHED_LEN + at] != 0u) return "table blob: check schedule counts do not match";
				}
				for (int i = 0; i < MB_SCHED_LEN; i++) {
					const uint32_t c = i == MB_SCHED_LEN - 1 ? 0u : cs[w * MB_SCHED_LEN + i], v = vs[w * MB_SCHED_LEN + i];
					if (c != 0u && ((c >> 25) == 0u || MB_CDESC_BASE(c) + 32u * MB_CDESC_DP(c) > (uint32_t)t.c_slots || MB_CDESC_DP(c) < 2u || MB_CDESC_DP(c) > MB_LDPC_DMAX ||
							MB_CDESC_TASK(c) >= (1u << MB_CDESC_LOG2S(c)) || MB_CDESC_GROUP(c) * 32u >= (uint32_t)t.P))
						return "table blob: check schedule out of range";
					if (v != 0u && ((v >> 24) == 0u || (v & 0xFFFFu) + 32u * ((v >> 16) & 0xFFu) > (uint32_t)t.v_slots || ((v >> 24) - 1u) * 32u + 32u > (uint32_t)t.N))
						return "table blob: variable schedule out of range";
				}
			}
			const uint8_t *cdeg = blob + t.off_cdeg;
			const uint32_t *cgb = reinterpret_cast<const uint32_t *>(blob + t.off_cgbase);
			const uint16_t *cos = reinterpret_cast<const uint16_t *>(blob + t.off_check_of_sorted), *voc = reinterpret_cast<const uint16_t *>(blob + t.off_var_of_cw);
			for (int c = 0; c < t.P; c++)
				if (cdeg[c] > MB_MAX_CDEG || cdeg[c] > cdeg[c & ~31] || mb_ldpc_cslot(cgb, cdeg[c & ~31], c | 31, cdeg[c & ~31] - 1) >= (uint32_t)t.c_slots || cos[c] >= t.P)
					return "table blob: check table out of range";
			for (int v = 0; v < t.N; v++)
				if (voc[v] >= t.N) return "table blob: variable map out of range";
		}
	}
	for (int c = 0; c < MB_NMODES; c++) {
		const MbMode &m = h.modes[c];
		size_t cells = (size_t)m.Nsymb * MB_NC;
		if (m.config != c || m.Nsymb <= 0 || m.Nsymb > MB_MAX_SYMB || m.rate_idx < 0 || m.rate_idx >= MB_NRATES || m.nBits > MB_N)
			return "table blob: bad mode record";
		if (!in(m.off_pinv, 4 * cells) || !in(m.off_pval, 4 * cells) || !in(m.off_invn, 4 * cells) ||
		    !in(m.off_pilot_cell, 2 * (size_t)m.nPilots) || !in(m.off_sym_cell, 2 * (size_t)m.nData) ||
		    !in(m.off_llr_dst, 2 * (size_t)m.nBits) || !in(m.off_llr_dst2, 2 * (size_t)m.nBits) || !in(m.off_const, 8 * (size_t)m.M) ||
		    !in(m.off_bit_var, 16 * (size_t)m.crc_bytes) || !in(m.off_scr, MB_N) || !in(m.off_crcbit, 16 * (size_t)m.crc_bytes) ||
		    !in(m.off_zf_src, 4 * (size_t)m.Nsymb * MB_ZF_STRIDE) || !in(m.off_pilot_rec, 16 * (size_t)m.nPilots) ||
		    !in(m.off_pilot_f, 8 * (size_t)m.nPilots) || (m.data_rec_words != 2 && m.data_rec_words != 4) ||
		    !in(m.off_data_rec, 4 * (size_t)m.data_rec_words * m.nData) || !in(m.off_virt, 4 * (size_t)m.nVirtual) ||
		    !in(m.off_pilot_neg, 8 * (size_t)m.Nsymb))
			return "table blob: mode table out of range";
	}
	return "";
}

// ---------------------------------------------------------------------------------------------------------------------
// ROBUST_0..2 (MFSK modes).  Their tables are NOT part of the blob (its layout and version are what other
// ranks import): they are derived from it on every handle and live in an extension region appended to the device copy, so the
// offsets below are relative to the same base as the blob's own (the decoder kernel addresses blob + off).
// Reference: common_defines.h:63-65, telecom_system.cc:2625-2645,2694-2700,1812-1817 (geometry), mfsk.cc:49-160 (tone plan).
// ---------------------------------------------------------------------------------------------------------------------
std::string mb_build_mfsk_ext(const std::vector<uint8_t> &blob, uint32_t base, MbMode modes[3], MbMfsk tones[3], std::vector<uint8_t> &ext)
{
	MbBlobHeader h;
	memcpy(&h, blob.data(), sizeof(h));
	Blob bl;
	static const int pre32[4] = {4, 20, 12, 28}, pre16[4] = {2, 10, 6, 14};
	static const int ack32[8] = {8, 14, 10, 24, 26, 2, 18, 30}, ack16[8] = {4, 7, 5, 12, 13, 1, 9, 15};
	static const int brk32[8] = {12, 28, 4, 6, 20, 16, 22, 30}, brk16[8] = {6, 14, 2, 3, 10, 8, 11, 15};
	for (int i = 0; i < 3; i++) {
		MbMode &m = modes[i];
		MbMfsk &t = tones[i];
		memset(&m, 0, sizeof(m));
		memset(&t, 0, sizeof(t));
		t.M = i == 0 ? 32 : 16, t.nStreams = i == 0 ? 1 : 2;
		t.nBits = i == 0 ? 5 : 4, t.tone_hop_step = i == 0 ? 13 : 7;
		const int global_offset = std::max(0, (MB_NC - t.nStreams * t.M) / 2);
		for (int k = 0; k < t.nStreams; k++) t.stream_offsets[k] = global_offset + k * t.M;
		for (int k = 0; k < 4; k++) t.preamble_tones[k] = i == 0 ? pre32[k] : pre16[k];
		for (int k = 0; k < 8; k++) t.ack_tones[k] = i == 0 ? ack32[k] : ack16[k], t.break_tones[k] = i == 0 ? brk32[k] : brk16[k];
		const int rate_num = i == 2 ? 4 : 1;
		m.config = 100 + i, m.M = 200 /* MOD_MFSK */, m.bps = t.nBits * t.nStreams, m.rate_num = rate_num, m.rate_idx = mb_rate_index(rate_num);
		const MbRate &r = h.rates[m.rate_idx];
		m.Nsymb = MB_N / m.bps, m.nData = m.Nsymb, m.nPilots = 0, m.nBits = MB_N;
		m.K = r.K, m.P = r.P, m.nReal = m.nBits - m.P, m.nVirtual = MB_N - m.nBits;
		m.frame_bytes = (m.nReal - 16) / 8, m.estimator = 1, m.phase_only = 0, m.preamble_nSymb = 4, m.boost = 1.33f;
		const uint16_t *var_of_cw = reinterpret_cast<const uint16_t *>(blob.data() + r.off_var_of_cw);
		// LLR i of cl_mfsk::demod (i = symbol * bps + stream * nBits + bit) -> bit de-interleaver (interleaver.cc:77-92, block nBits/10)
		// -> codeword position (no virtual bits, so the parity move of telecom_system.cc:1300-1308 is the identity) -> internal
		// variable -> hand-off slot
		const int bs = m.nBits / 10, nb = m.nBits / bs;
		std::vector<uint16_t> dst(m.nBits);
		for (int q = 0; q < m.nBits; q++) {
			const int cw = q < nb * bs ? (q % nb) * bs + q / nb : q;
			dst[q] = (uint16_t)MB_HANDOFF((unsigned)var_of_cw[cw]);
		}
		m.off_llr_dst = base + bl.put(dst);
		m.crc_bytes = m.nReal / 8;
		std::vector<uint16_t> bit_var(8 * m.crc_bytes);
		for (int q = 0; q < 8 * m.crc_bytes; q++) bit_var[q] = var_of_cw[q];
		m.off_bit_var = base + bl.put(bit_var);
		std::vector<uint8_t> scr(MB_N);
		uint32_t st[35];
		mb_srandom(st, 0);
		for (int q = 0; q < MB_N; q++) scr[q] = (uint8_t)(mb_random(st) % 2);
		m.off_scr = base + bl.put(scr);
		m.crc_reserved = 0;
		m.off_crcbit = base + bl.put(crc_bit_table(m.crc_bytes));
		uint16_t sreg = 0xFFFF;
		for (int k = 0; k < m.crc_bytes; k++) sreg = crc_zero_byte(sreg);
		m.crc_init = sreg;
	}
	bl.reserve(0, 256);
	ext.swap(bl.b);
	return "";
}
