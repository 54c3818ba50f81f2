// Evaluation metrics next to the output: Y-channel PSNR and SSIM of 8-bit frames with a cropped border, as the
// reference's evaluation driver computes them on the host from PNG files
// (CVSR_train/metric/psnr_ssim.py: calculate_psnr :278-316, _ssim :318-350, calculate_ssim :353-399, called with
// crop_border = 4 and test_y_channel = True on single-channel frames :447-478; to_y_channel :201-214 is the identity for one
// channel).  Doing it on the GPU keeps evaluation from bouncing every HR frame through the host and the file system.
//
//   PSNR = 20 log10(255 / sqrt(mean((a - b)^2)))          over the image without its `crop` border pixels
//   SSIM = mean over the "valid" region (another 5 pixels in) of
//          (2 mu_a mu_b + C1)(2 s_ab + C2) / ((mu_a^2 + mu_b^2 + C1)(s_a + s_b + C2))
//          with an 11 x 11 Gaussian window (sigma 1.5, cv2.getGaussianKernel), C1 = (0.01*255)^2, C2 = (0.03*255)^2, float64.
#include "common.cuh"

#define MT 16                 // tile edge (cropped-image pixels)
#define MH (MT + 10)          // with the 5-pixel window halo

struct MetricArgs {
    const unsigned char* a; const unsigned char* b;
    int H, W, crop, Hc, Wc, tiles_x, tiles_y;
    double g[11];             // normalised Gaussian taps
    double* partial;          // [B][tiles][2]: squared-error sum, SSIM sum
};

__global__ void __launch_bounds__(256) psnr_ssim_tile_kernel(const MetricArgs m) {
    __shared__ float A[MH][MH], Bv[MH][MH];
    __shared__ double Hs[5][MH][MT];          // horizontally filtered: a, b, a^2, b^2, ab
    __shared__ double red[2][8];
    const int tx0 = blockIdx.x * MT, ty0 = blockIdx.y * MT, img = blockIdx.z;
    const size_t base = (size_t)img * m.H * m.W;
    for (int e = threadIdx.x; e < MH * MH; e += 256) {
        const int r = e / MH, c = e - r * MH;
        // cropped coordinates, clamped into the cropped image (clamped pixels only feed window positions that are not valid)
        const int y = min(max(ty0 - 5 + r, 0), m.Hc - 1), x = min(max(tx0 - 5 + c, 0), m.Wc - 1);
        const size_t o = base + (size_t)(y + m.crop) * m.W + (x + m.crop);
        A[r][c] = (float)m.a[o];
        Bv[r][c] = (float)m.b[o];
    }
    __syncthreads();
    for (int e = threadIdx.x; e < MH * MT; e += 256) {
        const int r = e / MT, c = e - r * MT;
        double s0 = 0, s1 = 0, s2 = 0, s3 = 0, s4 = 0;
#pragma unroll
        for (int k = 0; k < 11; ++k) {
            const double av = A[r][c + k], bv = Bv[r][c + k], w = m.g[k];
            s0 += w * av; s1 += w * bv; s2 += w * av * av; s3 += w * bv * bv; s4 += w * av * bv;
        }
        Hs[0][r][c] = s0; Hs[1][r][c] = s1; Hs[2][r][c] = s2; Hs[3][r][c] = s3; Hs[4][r][c] = s4;
    }
    __syncthreads();
    const int ly = threadIdx.x >> 4, lx = threadIdx.x & 15;
    const int y = ty0 + ly, x = tx0 + lx;                       // cropped coordinates of this thread's pixel
    double se = 0.0, ss = 0.0;
    if (y < m.Hc && x < m.Wc) {
        const double d = (double)A[ly + 5][lx + 5] - (double)Bv[ly + 5][lx + 5];
        se = d * d;
        if (y >= 5 && y < m.Hc - 5 && x >= 5 && x < m.Wc - 5) {
            double v[5];
#pragma unroll
            for (int q = 0; q < 5; ++q) {
                double s = 0;
#pragma unroll
                for (int k = 0; k < 11; ++k) s += m.g[k] * Hs[q][ly + k][lx];
                v[q] = s;
            }
            const double C1 = (0.01 * 255) * (0.01 * 255), C2 = (0.03 * 255) * (0.03 * 255);
            const double mu1 = v[0], mu2 = v[1];
            const double s1 = v[2] - mu1 * mu1, s2 = v[3] - mu2 * mu2, s12 = v[4] - mu1 * mu2;
            ss = ((2 * mu1 * mu2 + C1) * (2 * s12 + C2)) / ((mu1 * mu1 + mu2 * mu2 + C1) * (s1 + s2 + C2));
        }
    }
    // deterministic block reduction: lanes (xor tree), then the 8 warps in fixed order
    for (int o = 16; o > 0; o >>= 1) {
        se += __shfl_xor_sync(0xffffffffu, se, o);
        ss += __shfl_xor_sync(0xffffffffu, ss, o);
    }
    if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = se; red[1][threadIdx.x >> 5] = ss; }
    __syncthreads();
    if (threadIdx.x < 2) {
        double s = 0;
        for (int k = 0; k < 8; ++k) s += red[threadIdx.x][k];
        const size_t t = ((size_t)img * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
        m.partial[t * 2 + threadIdx.x] = s;
    }
}

__global__ void psnr_ssim_final_kernel(const double* __restrict__ partial, int tiles, double npix, double nvalid,
                                       float* __restrict__ out) {
    __shared__ double red[2][32];
    const int img = blockIdx.x;
    double se = 0, ss = 0;
    for (int t = threadIdx.x; t < tiles; t += blockDim.x) {       // strided, then a fixed-order tree: deterministic
        se += partial[((size_t)img * tiles + t) * 2];
        ss += partial[((size_t)img * tiles + t) * 2 + 1];
    }
    for (int o = 16; o > 0; o >>= 1) {
        se += __shfl_xor_sync(0xffffffffu, se, o);
        ss += __shfl_xor_sync(0xffffffffu, ss, o);
    }
    if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = se; red[1][threadIdx.x >> 5] = ss; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0, b = 0;
        for (int k = 0; k < (int)(blockDim.x >> 5); ++k) { a += red[0][k]; b += red[1][k]; }
        const double mse = a / npix;
        out[img * 2] = mse == 0.0 ? INFINITY : (float)(20.0 * log10(255.0 / sqrt(mse)));
        out[img * 2 + 1] = nvalid > 0 ? (float)(b / nvalid) : 0.f;
    }
}

// a, b: [B,H,W] uint8 frames (device); out: [B][2] floats = (PSNR in dB, SSIM); scratch: B * ceil(Hc/16) * ceil(Wc/16) * 2 doubles
// with Hc = H - 2 crop, Wc = W - 2 crop.  Needs Hc, Wc >= 11 (one valid SSIM window).
extern "C" int fcvsr_psnr_ssim_u8(const unsigned char* a, const unsigned char* b, int B, int H, int W, int crop, double* scratch,
                                  float* out, cudaStream_t st) {
    if (!a || !b || !scratch || !out || B <= 0 || crop < 0) return FCVSR_ERR_ARG;
    MetricArgs m;
    m.a = a; m.b = b; m.H = H; m.W = W; m.crop = crop; m.Hc = H - 2 * crop; m.Wc = W - 2 * crop;
    if (m.Hc < 11 || m.Wc < 11) return FCVSR_ERR_ARG;
    m.tiles_x = (m.Wc + MT - 1) / MT; m.tiles_y = (m.Hc + MT - 1) / MT;
    double s = 0;
    for (int k = 0; k < 11; ++k) { m.g[k] = exp(-(double)((k - 5) * (k - 5)) / (2.0 * 1.5 * 1.5)); s += m.g[k]; }
    for (int k = 0; k < 11; ++k) m.g[k] /= s;
    m.partial = scratch;
    psnr_ssim_tile_kernel<<<dim3(m.tiles_x, m.tiles_y, B), 256, 0, st>>>(m);
    psnr_ssim_final_kernel<<<B, 256, 0, st>>>(scratch, m.tiles_x * m.tiles_y, (double)m.Hc * m.Wc,
                                              (double)(m.Hc - 10) * (m.Wc - 10), out);
    return fcvsr_launch_status();
}
