/* Stand-in for the author's TGM "FFT.h" (class spectrum), which is NOT in /root/reference and has no pinned version
 * (PARITY UNPINNED, see oracle/jade_oracle.h).  Forwards to the oracle's float32 FFT so that the compiled reference
 * and the restated oracle use the very same FFT and can be compared bit for bit.  jo_power_f32 keeps its plan per thread and
 * per size (rebuilt only when the size changes, as after setFFTSize, Spectrogram.cpp:215), so timing the compiled reference
 * through this shim does not pay for re-planning on every frame.  TEST INFRASTRUCTURE ONLY. */
#pragma once
#include <cstddef>
#include <vector>
extern "C" int jo_power_f32(const float* x, int n, float* power);
class spectrum
{
public:
    explicit spectrum(int n = 1024) : m_n(n) {}
    void setFFTSize(size_t n) { m_n = int(n); }
    void power(float* in, std::vector<float>& out) /* call site Spectrogram.cpp:144 */
    {
        if (out.size() < size_t(m_n / 2 + 1)) out.resize(size_t(m_n / 2 + 1));
        jo_power_f32(in, m_n, out.data());
    }
private:
    int m_n;
};
