// ACAP per-vertex rotation / shear (pyACAP.GetRS with _R = 1) on the GPU -- SURVEY.md 8f-3.
//
// Replaces the CPU / OpenMP / Eigen / OpenMesh path of the reference's ACAP library
// (ACAP/pyACAPv1.zip: src/FeatureVector.cpp RefMesh::RefMesh :81-173, RefMesh::GetRS :428-590,
// src/Align.cpp AffineAlign :60-100, polarDec :31-56), which edittool/__init__.py:109 calls once per
// deformed mesh.  Arithmetic is float64 like the reference; R and S leave as float32 for the deform kernel.
//
//   rest (once per mesh)   : one-ring fans (host, gm_acap_build_rings), then per vertex the fourth-root
//                            cotangent weights, the unit normal and AtA^-1 = (sum p p^T)^-1
//   per deformation        : vertex normals of the deformed mesh, then per vertex
//                            T = (AtA^-1 sum_k p_k v_k^T)^T,  T = r s (polar),  R = r^T,  S = s
// The polar factors come from the eigen-decomposition of T^T T (cyclic Jacobi): s = V diag(sigma') V^T,
// r = T V diag(1/sigma') V^T, where sigma' carries a minus sign on the smallest singular value when det T < 0
// -- exactly where Align.cpp:38-53 moves the reflection.
#include <algorithm>
#include <cmath>
#include <vector>

#include "common.cuh"
#include "kernels.h"

namespace gm {

namespace {

constexpr int kThreads = 128;
constexpr double kNormalScale = 0.3;   // RefMesh::normalScale, FeatureVector.cpp:22
constexpr double kEps = 1e-10;         // Align.h:10

struct D3 { double x, y, z; };
__device__ __forceinline__ D3 ld3(const double* p, int i) { return {p[3 * i], p[3 * i + 1], p[3 * i + 2]}; }
__device__ __forceinline__ D3 sub(D3 a, D3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
__device__ __forceinline__ D3 scl(D3 a, double s) { return {a.x * s, a.y * s, a.z * s}; }
__device__ __forceinline__ double dot(D3 a, D3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ double nrm(D3 a) { return sqrt(dot(a, a)); }
__device__ __forceinline__ D3 cross(D3 a, D3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }

// OpenMesh update_normals(): unit face normals summed over the incident faces, then normalised.
__global__ void __launch_bounds__(kThreads)
acap_normals_kernel(int Vn, const double* __restrict__ V, const int* __restrict__ F, const int* __restrict__ face_off,
                    const int* __restrict__ face_list, double* __restrict__ normals)
{
	pdl_sync();
	const int i = blockIdx.x * kThreads + threadIdx.x;
	if (i >= Vn)
		return;
	D3 n = {0, 0, 0};
	for (int e = face_off[i]; e < face_off[i + 1]; e++) {
		const int f = face_list[e];
		const D3 a = ld3(V, F[3 * f]), b = ld3(V, F[3 * f + 1]), c = ld3(V, F[3 * f + 2]);
		D3 fn = cross(sub(b, a), sub(c, a));
		const double l = nrm(fn);
		if (l > 0) n = {n.x + fn.x / l, n.y + fn.y / l, n.z + fn.z / l};
	}
	const double l = nrm(n);
	if (l > 0) n = scl(n, 1.0 / l);
	normals[3 * i] = n.x; normals[3 * i + 1] = n.y; normals[3 * i + 2] = n.z;
}

// FeatureVector.cpp:33-39
__device__ __forceinline__ double cotan(D3 a, D3 b)
{
	const double na = nrm(a), nb = nrm(b);
	if (na < kEps || nb < kEps) return 0;
	const double c = dot(a, b) / (na * nb);
	if (c == 1) return 1;
	return c / sqrt(1 - c * c);
}

__device__ __forceinline__ void add_outer(double (&A)[9], D3 p, D3 v)
{
	A[0] += p.x * v.x; A[1] += p.x * v.y; A[2] += p.x * v.z;
	A[3] += p.y * v.x; A[4] += p.y * v.y; A[5] += p.y * v.z;
	A[6] += p.z * v.x; A[7] += p.z * v.y; A[8] += p.z * v.z;
}

// RefMesh::RefMesh, FeatureVector.cpp:95-172
__global__ void __launch_bounds__(kThreads)
acap_rest_kernel(int Vn, const double* __restrict__ V, const int* __restrict__ ring_off, const int* __restrict__ ring,
                 const double* __restrict__ normals, double* __restrict__ sqrt_w, double* __restrict__ ata_inv)
{
	pdl_sync();
	const int i = blockIdx.x * kThreads + threadIdx.x;
	if (i >= Vn)
		return;
	const int b = ring_off[i], n = ring_off[i + 1] - b;
	double A[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
	const D3 p = ld3(V, i);
	double lens = 0;
	for (int k = 0; k < n; k++) {
		const D3 cur = ld3(V, ring[b + k]);
		const D3 prv = ld3(V, ring[b + (k + n - 1) % n]), nxt = ld3(V, ring[b + (k + 1) % n]);
		const double w1 = cotan(sub(p, prv), sub(cur, prv));
		const double w2 = cotan(sub(p, nxt), sub(cur, nxt));
		const double x = 0.5 * (w1 + w2);
		double w = sqrt(x <= 0 ? exp(x) : 1 + x);                 // sqrt(sexp(.)), :134, sexp :69-72
		if (w != w || w > 100000) w = 1;                          // :136-141
		const double sw = sqrt(w);                                // the edge vector is scaled by sqrt(w), :155
		sqrt_w[b + k] = sw;
		const D3 q = sub(cur, p);
		lens += nrm(q);
		const D3 pk = scl(q, sw);
		add_outer(A, pk, pk);
	}
	double* out = ata_inv + 9 * (size_t)i;
	if (n == 0) {
		for (int e = 0; e < 9; e++) out[e] = (e % 4 == 0) ? 1.0 : 0.0;
		return;
	}
	const D3 pn = scl(ld3(normals, i), lens / n * kNormalScale);   // :164
	add_outer(A, pn, pn);
	// 3x3 inverse (Align.cpp:75)
	const double c00 = A[4] * A[8] - A[5] * A[7], c01 = A[5] * A[6] - A[3] * A[8], c02 = A[3] * A[7] - A[4] * A[6];
	const double det = A[0] * c00 + A[1] * c01 + A[2] * c02;
	const double id = 1.0 / det;
	out[0] = c00 * id; out[1] = (A[2] * A[7] - A[1] * A[8]) * id; out[2] = (A[1] * A[5] - A[2] * A[4]) * id;
	out[3] = c01 * id; out[4] = (A[0] * A[8] - A[2] * A[6]) * id; out[5] = (A[2] * A[3] - A[0] * A[5]) * id;
	out[6] = c02 * id; out[7] = (A[1] * A[6] - A[0] * A[7]) * id; out[8] = (A[0] * A[4] - A[1] * A[3]) * id;
}

// Cyclic Jacobi eigen-decomposition of a symmetric 3x3 (row-major a[9]); eigenvectors in the columns of v.
__device__ void jacobi_eig3(double (&a)[9], double (&v)[9])
{
	for (int e = 0; e < 9; e++) v[e] = (e % 4 == 0) ? 1.0 : 0.0;
	for (int sweep = 0; sweep < 30; sweep++) {
		const double off = a[1] * a[1] + a[2] * a[2] + a[5] * a[5];
		const double diag = a[0] * a[0] + a[4] * a[4] + a[8] * a[8];
		if (off <= 1e-32 * diag || off == 0)
			break;
		for (int pq = 0; pq < 3; pq++) {
			const int p = (pq == 2) ? 1 : 0, q = (pq == 0) ? 1 : 2;
			const double apq = a[3 * p + q];
			if (apq == 0)
				continue;
			const double theta = (a[3 * q + q] - a[3 * p + p]) / (2 * apq);
			const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1));
			const double c = 1 / sqrt(t * t + 1), s = t * c;
			for (int k = 0; k < 3; k++) {          // A <- A J
				const double akp = a[3 * k + p], akq = a[3 * k + q];
				a[3 * k + p] = c * akp - s * akq;
				a[3 * k + q] = s * akp + c * akq;
			}
			for (int k = 0; k < 3; k++) {          // A <- J^T A
				const double apk = a[3 * p + k], aqk = a[3 * q + k];
				a[3 * p + k] = c * apk - s * aqk;
				a[3 * q + k] = s * apk + c * aqk;
			}
			for (int k = 0; k < 3; k++) {          // V <- V J
				const double vkp = v[3 * k + p], vkq = v[3 * k + q];
				v[3 * k + p] = c * vkp - s * vkq;
				v[3 * k + q] = s * vkp + c * vkq;
			}
		}
	}
}

// RefMesh::GetRS, FeatureVector.cpp:447-494 + output :560-590
__global__ void __launch_bounds__(kThreads)
acap_rs_kernel(int Vn, const double* __restrict__ V0, const double* __restrict__ V1, const int* __restrict__ ring_off,
               const int* __restrict__ ring, const double* __restrict__ sqrt_w, const double* __restrict__ n0,
               const double* __restrict__ n1, const double* __restrict__ ata_inv, float* __restrict__ R_out,
               float* __restrict__ S_out)
{
	pdl_sync();
	const int i = blockIdx.x * kThreads + threadIdx.x;
	if (i >= Vn)
		return;
	const int b = ring_off[i], n = ring_off[i + 1] - b;
	float* Ro = R_out + 9 * (size_t)i;
	float* So = S_out + 9 * (size_t)i;
	if (n == 0) {                                                  // :466-469
		for (int e = 0; e < 9; e++) Ro[e] = So[e] = (e % 4 == 0) ? 1.0f : 0.0f;
		return;
	}
	const D3 p0 = ld3(V0, i), p1 = ld3(V1, i);
	double M[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};                     // sum_k p_k v_k^T
	double len0 = 0, len1 = 0;
	for (int k = 0; k < n; k++) {
		const int j = ring[b + k];
		const double sw = sqrt_w[b + k];
		const D3 q0 = sub(ld3(V0, j), p0), q1 = sub(ld3(V1, j), p1);
		len0 += nrm(q0);
		len1 += nrm(q1);
		add_outer(M, scl(q0, sw), scl(q1, sw));
	}
	add_outer(M, scl(ld3(n0, i), len0 / n * kNormalScale), scl(ld3(n1, i), len1 / n * kNormalScale));
	// T = (AtA^-1 M)^T  (Align.cpp:77-95)
	const double* Ai = ata_inv + 9 * (size_t)i;
	double T[9];
	for (int r = 0; r < 3; r++)
		for (int c = 0; c < 3; c++)
			T[3 * c + r] = Ai[3 * r] * M[c] + Ai[3 * r + 1] * M[3 + c] + Ai[3 * r + 2] * M[6 + c];
	// polar decomposition (Align.cpp:31-56)
	double G[9], Vv[9];
	for (int r = 0; r < 3; r++)
		for (int c = 0; c < 3; c++)
			G[3 * r + c] = T[r] * T[c] + T[3 + r] * T[3 + c] + T[6 + r] * T[6 + c];      // T^T T
	jacobi_eig3(G, Vv);
	double sig[3] = {sqrt(fmax(G[0], 0.0)), sqrt(fmax(G[4], 0.0)), sqrt(fmax(G[8], 0.0))};
	const double detT = T[0] * (T[4] * T[8] - T[5] * T[7]) - T[1] * (T[3] * T[8] - T[5] * T[6]) + T[2] * (T[3] * T[7] - T[4] * T[6]);
	if (detT < 0) {
		int m = 0;
		if (sig[1] < sig[m]) m = 1;
		if (sig[2] < sig[m]) m = 2;
		sig[m] = -sig[m];
	}
	// s = V diag(sig) V^T ;  r = T V diag(1/sig) V^T
	double s[9], w[9];
	const double big = fmax(fabs(sig[0]), fmax(fabs(sig[1]), fabs(sig[2])));
	for (int r = 0; r < 3; r++)
		for (int c = 0; c < 3; c++) {
			double acc = 0, acw = 0;
			for (int k = 0; k < 3; k++) {
				acc += Vv[3 * r + k] * sig[k] * Vv[3 * c + k];
				// a (numerically) zero singular value carries no rotation information; keep r finite
				const double inv = fabs(sig[k]) > 1e-14 * big ? 1.0 / sig[k] : 0.0;
				acw += Vv[3 * r + k] * inv * Vv[3 * c + k];
			}
			s[3 * r + c] = acc;
			w[3 * r + c] = acw;
		}
	for (int r = 0; r < 3; r++)
		for (int c = 0; c < 3; c++) {
			const double rv = T[3 * r] * w[c] + T[3 * r + 1] * w[3 + c] + T[3 * r + 2] * w[6 + c];
			Ro[3 * c + r] = (float)rv;               // R = r^T (:562-570)
			So[3 * r + c] = (float)s[3 * r + c];
		}
}

} // namespace

int launch_acap_rest(int Vn, const double* V, const int* F, const int* ring_off, const int* ring, const int* face_off,
                     const int* face_list, double* sqrt_w, double* normals, double* ata_inv, cudaStream_t stream)
{
	if (Vn <= 0) return GM_OK;
	const int blocks = (Vn + kThreads - 1) / kThreads;
	launch_k(acap_normals_kernel, dim3(blocks), dim3(kThreads), 0, stream, Vn, V, F, face_off, face_list, normals);
	launch_k(acap_rest_kernel, dim3(blocks), dim3(kThreads), 0, stream, Vn, V, ring_off, ring, normals, sqrt_w, ata_inv);
	return GM_OK;
}

int launch_acap_get_rs(int Vn, const double* V0, const double* V1, const int* F, const int* ring_off, const int* ring,
                       const int* face_off, const int* face_list, const double* sqrt_w, const double* n0,
                       const double* ata_inv, double* n1_scratch, float* R_out, float* S_out, cudaStream_t stream)
{
	if (Vn <= 0) return GM_OK;
	const int blocks = (Vn + kThreads - 1) / kThreads;
	launch_k(acap_normals_kernel, dim3(blocks), dim3(kThreads), 0, stream, Vn, V1, F, face_off, face_list, n1_scratch);
	launch_k(acap_rs_kernel, dim3(blocks), dim3(kThreads), 0, stream, Vn, V0, V1, ring_off, ring, sqrt_w, n0, n1_scratch, ata_inv, R_out, S_out);
	return GM_OK;
}

// Host: one-ring fans in cyclic order and the vertex -> incident-face lists.  For every vertex v each incident face
// (v, a, b) gives a spoke a -> b; closed fans chain into a cycle, boundary fans into a path that starts at the
// neighbour no spoke points to (OpenMesh hands boundary rings out from the boundary edge as well).
int acap_build_rings_host(int Vn, int Fn, const int* F, int* ring_off, int* ring, int* face_off, int* face_list)
{
	std::vector<int> deg(Vn + 1, 0);
	for (int f = 0; f < Fn; f++)
		for (int k = 0; k < 3; k++) {
			const int v = F[3 * f + k];
			if (v < 0 || v >= Vn) return GM_ERR_BAD_ARGUMENT;
			deg[v + 1]++;
		}
	face_off[0] = 0;
	for (int v = 0; v < Vn; v++) face_off[v + 1] = face_off[v] + deg[v + 1];
	std::vector<int> fill(face_off, face_off + Vn);
	std::vector<int> spoke_a(3 * (size_t)Fn), spoke_b(3 * (size_t)Fn);
	for (int f = 0; f < Fn; f++)
		for (int k = 0; k < 3; k++) {
			const int v = F[3 * f + k];
			const int at = fill[v]++;
			face_list[at] = f;
			spoke_a[at] = F[3 * f + (k + 1) % 3];
			spoke_b[at] = F[3 * f + (k + 2) % 3];
		}
	int out = 0;
	ring_off[0] = 0;
	for (int v = 0; v < Vn; v++) {
		const int s0 = face_off[v], s1 = face_off[v + 1];
		if (s1 > s0) {
			int start = spoke_a[s0];
			for (int e = s0; e < s1; e++) {             // a chain head: no spoke points to it
				bool pointed = false;
				for (int g = s0; g < s1 && !pointed; g++) pointed = (spoke_b[g] == spoke_a[e]);
				if (!pointed) { start = spoke_a[e]; break; }
			}
			int cur = start;
			const int first_out = out;
			for (int steps = 0; steps <= s1 - s0; steps++) {
				bool seen = false;
				for (int r = first_out; r < out && !seen; r++) seen = (ring[r] == cur);
				if (seen) break;
				ring[out++] = cur;
				int next = -1;
				for (int e = s0; e < s1; e++)
					if (spoke_a[e] == cur) { next = spoke_b[e]; break; }
				if (next < 0) break;
				cur = next;
			}
		}
		ring_off[v + 1] = out;
	}
	return GM_OK;
}

} // namespace gm
