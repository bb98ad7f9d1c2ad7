/* tests/csrc/m3dtest.cu — TEST-ONLY kernels (never linked into the product).
 *
 * Exhaustive proof over all 2^32 float bit patterns that three formulations of the reference's angle gate
 * (lesson_16.cu:666-671) agree:
 *   gate_ref       the upstream expression, typed the way upstream types it
 *   m3d::angle_gate the product's device function (csrc/m3dreg_kernels.cuh)
 *   gate_restated  the branchy restatement used by the CPU oracle (oracle/m3d_oracle.c: orc_angle_gate)
 */
#include "../../mandala-mapping_b200/csrc/m3dreg_kernels.cuh"
#include <cstdio>

__device__ __forceinline__ bool gate_ref(float dotProduct)
{
	float angle = acos(dotProduct);
	float angled = angle * 180.0f / M_PI;
	if (angled < 0) angled = -angled;
	return angled < 90.0f;
}

__device__ __forceinline__ bool gate_restated(float d)
{
	float a = fabsf(d);
	if (!(a <= 1.0f)) return false;
	if (a > __uint_as_float(0x3F0F5C29u)) return d > 0.0f;
	float t2 = __fmul_rn(d, d);
	float p = __fmaf_rn(t2, __uint_as_float(0x3D10ECEFu), __uint_as_float(0x3C8B1ABBu));
	p = __fmaf_rn(p, t2, __uint_as_float(0x3CFC028Cu));
	p = __fmaf_rn(p, t2, __uint_as_float(0x3D372139u));
	p = __fmaf_rn(p, t2, __uint_as_float(0x3D9993DBu));
	p = __fmaf_rn(p, t2, __uint_as_float(0x3E2AAAC6u));
	float q = __fmul_rn(t2, p);
	float s = __fmaf_rn(q, d, d);
	float ang = __fmaf_rn(__uint_as_float(0x3F6EE581u), __uint_as_float(0x3FD774EBu), -s);
	float deg = __fmul_rn(ang, 180.0f);
	float angled = (float)((double)deg / 3.14159265358979323846);
	if (angled < 0) angled = -angled;
	return angled < 90.0f;
}

/* out[0] = mismatches(ref vs product), out[1] = mismatches(ref vs restated), out[2] = #accepted,
 * out[3] = min accepted bit pattern, out[4] = max accepted bit pattern, out[5] = first mismatching pattern + 1 */
__global__ void k_gate_exhaustive(unsigned long long *out)
{
	unsigned long long mm1 = 0, mm2 = 0, acc = 0, mn = 0xFFFFFFFFull, mx = 0, bad = 0;
	unsigned long long total = 1ull << 32;
	for (unsigned long long u = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; u < total; u += (unsigned long long)gridDim.x * blockDim.x) {
		float d = __uint_as_float((uint32_t)u);
		bool r = gate_ref(d), p = m3d::angle_gate(d), s = gate_restated(d);
		if (r != p) { mm1++; if (!bad) bad = u + 1; }
		if (r != s) { mm2++; if (!bad) bad = u + 1; }
		if (r) { acc++; if (u < mn) mn = u; if (u > mx) mx = u; }
	}
	atomicAdd(&out[0], mm1);
	atomicAdd(&out[1], mm2);
	atomicAdd(&out[2], acc);
	atomicMin(&out[3], mn);
	atomicMax(&out[4], mx);
	if (bad) atomicMax(&out[5], bad);
}

__global__ void k_gate_window(uint32_t lo, uint32_t count, unsigned char *out)
{
	uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < count) out[i] = gate_ref(__uint_as_float(lo + i)) ? 1 : 0;
}

extern "C" int m3dtest_gate_exhaustive(unsigned long long *out6)
{
	unsigned long long *d = 0;
	unsigned long long init[6] = {0, 0, 0, 0xFFFFFFFFull, 0, 0};
	cudaError_t e = cudaMalloc((void **)&d, sizeof(init));
	if (e != cudaSuccess) return (int)e;
	cudaMemcpy(d, init, sizeof(init), cudaMemcpyHostToDevice);
	k_gate_exhaustive<<<148 * 16, 256>>>(d);
	e = cudaDeviceSynchronize();
	if (e == cudaSuccess) e = cudaMemcpy(out6, d, sizeof(init), cudaMemcpyDeviceToHost);
	cudaFree(d);
	return (int)e;
}

extern "C" int m3dtest_gate_window(uint32_t lo, uint32_t count, unsigned char *out_host)
{
	unsigned char *d = 0;
	cudaError_t e = cudaMalloc((void **)&d, count);
	if (e != cudaSuccess) return (int)e;
	k_gate_window<<<(count + 255) / 256, 256>>>(lo, count, d);
	e = cudaDeviceSynchronize();
	if (e == cudaSuccess) e = cudaMemcpy(out_host, d, count, cudaMemcpyDeviceToHost);
	cudaFree(d);
	return (int)e;
}
