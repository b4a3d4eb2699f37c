/*
 * Spectrogram.h -- drop-in replacement for the reference's `class Spectrogram` (Spectrogram.h:81-169) whose hot loop
 * (Spectrogram.cpp:37-135: framing, window, FFT power, channel mix, dB, ring) runs on a B200 through libjade_gpu.so.
 *
 * The public surface is the reference's: nested enums ChannelMixMode / Windows / FeedPercentage, the setters
 * (setSamplerate, setchannels, setFFTSize, setclosestFFTSize_ms, setmemoryTime_s, setfeed_percent, setPauseMode,
 * setWindow), getnextpowerof2, getSpectrumSize, getMemorySize, getMem, getSamplerate and
 * processSynchronBlock(std::vector<std::vector<float>>&, juce::MidiBuffer&).  Header-only; link with -ljade_gpu.
 *
 * With JUCE and the author's TGM library on the include path the class derives from the real SynchronBlockProcessor
 * exactly like the reference.  Without them (tests, offline tools) a minimal re-blocker with the same three calls the
 * plugin uses (preparetoProcess, setDesiredBlockSizeSamples, processBlock; PluginProcessor.cpp:108,148,
 * Spectrogram.cpp:164) and an empty juce::MidiBuffer stand in.
 *
 * There is no CPU implementation behind this class: without a CUDA device every processing call returns -1 and
 * lastError() says why (the reference's int-status convention; no exceptions cross the audio thread).
 */
#pragma once

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "CColorpalette.h"
#include "jade_gpu.h"

#if defined(__has_include)
#if __has_include("SynchronBlockProcessor.h") && __has_include(<juce_audio_processors/juce_audio_processors.h>)
#define JADE_HAVE_TGM_JUCE 1
#endif
#endif

#if defined(JADE_HAVE_TGM_JUCE)
#include "SynchronBlockProcessor.h"
#else
namespace juce
{
class MidiBuffer
{
};
} // namespace juce

/* Host block size -> fixed analysis block re-blocker (stand-in for the TGM SynchronBlockProcessor).  Audio passes
 * through untouched; complete blocks of `desired` samples per channel are handed to processSynchronBlock. */
class SynchronBlockProcessor
{
public:
	SynchronBlockProcessor() {}
	virtual ~SynchronBlockProcessor() {}
	void preparetoProcess(int channels, int maxBlockSize)
	{
		(void)maxBlockSize;
		m_NrOfChannels = channels > 0 ? channels : 1;
		resetAccumulator();
	}
	void setDesiredBlockSizeSamples(int n)
	{
		m_desired = n > 0 ? n : 1;
		resetAccumulator();
	}
	virtual int processSynchronBlock(std::vector<std::vector<float>>&, juce::MidiBuffer&) = 0;

	/* planar channel pointers, any numSamples */
	int processBlock(const float* const* channelData, int numChannels, int numSamples, juce::MidiBuffer& midi)
	{
		int rc = 0;
		int done = 0;
		while (done < numSamples) {
			const int take = std::min(numSamples - done, m_desired - m_fill);
			for (int ch = 0; ch < m_NrOfChannels; ++ch) {
				const float* src = channelData[ch < numChannels ? ch : numChannels - 1] + done;
				std::copy(src, src + take, m_acc[size_t(ch)].begin() + m_fill);
			}
			m_fill += take;
			done += take;
			if (m_fill == m_desired) {
				const int r = processSynchronBlock(m_acc, midi);
				if (r != 0)
					rc = r;
				m_fill = 0;
			}
		}
		return rc;
	}
	/* juce::AudioBuffer<float>-shaped argument (getNumChannels / getNumSamples / getReadPointer) */
	template <typename AudioBufferLike>
	int processBlock(AudioBufferLike& buffer, juce::MidiBuffer& midi)
	{
		std::vector<const float*> ptr(size_t(buffer.getNumChannels()));
		for (int ch = 0; ch < buffer.getNumChannels(); ++ch)
			ptr[size_t(ch)] = buffer.getReadPointer(ch);
		return processBlock(ptr.data(), buffer.getNumChannels(), buffer.getNumSamples(), midi);
	}

protected:
	void resetAccumulator()
	{
		m_acc.assign(size_t(m_NrOfChannels), std::vector<float>(size_t(m_desired), 0.f));
		m_fill = 0;
	}
	int m_NrOfChannels = 2;
	int m_desired = 1024;
	int m_fill = 0;
	std::vector<std::vector<float>> m_acc;
};
#endif

class Spectrogram : public SynchronBlockProcessor
{
public:
	enum class ChannelMixMode
	{
		AbsMean,
		Max,
		Min,
		Left,
		Right
	};
	enum class Windows
	{
		Rect,
		Hann,
		Hamming,
		BlackmanHarris,
		FlatTop,
		HannPoisson
	};
	enum class FeedPercentage
	{
		perc100,
		perc50,
		perc25,
		perc10
	};

	/* Defaults of the reference constructor (Spectrogram.cpp:16-24): 48 kHz, 2 channels, FFT 1024, 100 % feed,
	 * 1 s memory, Hann, AbsMean.  JADE_DEVICE selects the GPU (default 0). */
	Spectrogram() : SynchronBlockProcessor()
	{
		const char* dev = std::getenv("JADE_DEVICE");
		if (jade_create(dev ? std::atoi(dev) : 0, &m_engine) != JADE_OK) {
			m_error = jade_last_error(nullptr);
			std::fprintf(stderr, "Spectrogram: %s\n", m_error.c_str());
			m_engine = nullptr;
		}
		buildmem();
	}
	virtual ~Spectrogram()
	{
		if (m_engine)
			jade_destroy(m_engine);
	}
	Spectrogram(const Spectrogram&) = delete;
	Spectrogram& operator=(const Spectrogram&) = delete;

	/* Spectrogram.cpp:37-135.  data[channel][fftsize]; returns 0 (the reference always does) or -1 on error. */
	virtual int processSynchronBlock(std::vector<std::vector<float>>& data, juce::MidiBuffer& midiMessages)
	{
		(void)midiMessages;
		if (!m_engine || !m_configured)
			return -1;
		if (data.size() < m_channels)
			return setError("processSynchronBlock: fewer input channels than setchannels()"); /* reference: out of bounds */
		const float* ptr[16];
		for (size_t cc = 0; cc < m_channels; ++cc) {
			if (data[cc].size() < m_fftsize)
				return setError("processSynchronBlock: block shorter than the FFT size");
			ptr[cc] = data[cc].data();
		}
		if (jade_push_samples(m_engine, ptr, int(m_channels), int(m_fftsize)) != JADE_OK)
			return setError(jade_last_error(m_engine));
		if (!m_PauseMode) /* :111-118 */
			m_newEntryCounter += m_feedblocks;
		return 0;
	}

	// setter (Spectrogram.cpp:148-211)
	void setSamplerate(float samplerate)
	{
		m_fs = samplerate;
		buildmem();
	}
	void setchannels(size_t newchannels)
	{
		m_channels = newchannels;
		buildmem();
	}
	void setFFTSize(size_t newFFTSize)
	{
		m_fftsize = newFFTSize;
		setDesiredBlockSizeSamples(int(m_fftsize));
		buildmem();
		m_newEntryCounter = kAllNew;
	}
	void setclosestFFTSize_ms(float fftsize_ms)
	{
		m_fftsize = getnextpowerof2(fftsize_ms);
		setDesiredBlockSizeSamples(int(m_fftsize));
		buildmem();
	}
	void setmemoryTime_s(float memsize_s)
	{
		m_memsize_s = memsize_s;
		buildmem();
	}
	void setfeed_percent(FeedPercentage feed)
	{
		switch (feed) {
		case FeedPercentage::perc100: m_feed_percent = 100; break;
		case FeedPercentage::perc50: m_feed_percent = 50; break;
		case FeedPercentage::perc25: m_feed_percent = 25; break;
		case FeedPercentage::perc10: m_feed_percent = 10; break;
		}
		buildmem();
	}
	void setPauseMode(bool mode)
	{
		m_PauseMode = mode;
		if (m_engine)
			jade_set_pause(m_engine, mode ? 1 : 0);
	}
	void setWindow(Spectrogram::Windows win) /* table only, no buffer reset (Spectrogram.h:123) */
	{
		m_windowChoice = win;
		if (m_engine && m_configured)
			jade_set_window(m_engine, int(win));
	}
	/* extension: the reference fixes m_mode = AbsMean (Spectrogram.cpp:21) and has no setter */
	void setMixMode(ChannelMixMode mode)
	{
		m_mode = mode;
		buildmem();
	}

	size_t getnextpowerof2(float fftsize_ms) /* Spectrogram.cpp:171-176 */
	{
		float firstguessFFTSize = float(fftsize_ms * 0.001 * m_fs);
		int nextpowerof2 = int(std::log(double(firstguessFFTSize)) / std::log(double(2.f))) + 1;
		return size_t(std::pow(double(2.f), nextpowerof2));
	}

	int getSpectrumSize() { return int(m_freqsize); }
	int getMemorySize() { return m_memsize_blocks; }
	float getSamplerate() { return m_fs; }

	/* Spectrogram.cpp:295-331: copies the columns written since the previous call into the SAME ring indices of `mem`
	 * (mem mirrors the ring, it is not time-ordered), the whole ring if at least a ring's worth is new; returns the
	 * number of new columns (the reference's counter, "everything" = 1215752192 after a rebuild) and the ring write
	 * position in pos; -1 if mem has a different number of columns. */
	int getMem(std::vector<std::vector<float>>& mem, int& pos)
	{
		if (!m_engine || !m_configured)
			return -1;
		const int W = m_memsize_blocks, B = int(m_freqsize);
		if (mem.size() != size_t(W))
			return -1;
		int64_t total = 0;
		jade_ring_info(m_engine, nullptr, nullptr, nullptr, &total);
		if (size_t(m_newEntryCounter) >= mem.size()) {
			m_scratch.resize(size_t(W) * B);
			if (jade_read_ring_db(m_engine, m_scratch.data()) != JADE_OK)
				return setError(jade_last_error(m_engine));
			for (int kk = 0; kk < W; ++kk)
				std::copy(m_scratch.begin() + size_t(kk) * B, m_scratch.begin() + size_t(kk + 1) * B, mem[size_t(kk)].begin());
			int n = 0;
			int64_t first = 0;
			jade_fetch_columns(m_engine, nullptr, nullptr, 0, &n, &first); /* mark everything as seen */
		} else if (m_newEntryCounter > 0) {
			m_scratch.resize(size_t(m_newEntryCounter) * B);
			int n = 0;
			int64_t first = 0;
			if (jade_fetch_columns(m_engine, nullptr, m_scratch.data(), m_newEntryCounter, &n, &first) != JADE_OK)
				return setError(jade_last_error(m_engine));
			for (int i = 0; i < n; ++i) {
				const size_t slot = size_t((first + i) % W);
				std::copy(m_scratch.begin() + size_t(i) * B, m_scratch.begin() + size_t(i + 1) * B, mem[slot].begin());
			}
		}
		const int newVals = m_newEntryCounter;
		m_newEntryCounter = 0;
		pos = int(total % W);
		return newVals;
	}

	// ---- extensions over the reference ----
	bool ok() const { return m_engine != nullptr && m_configured; }
	const std::string& lastError() const { return m_error; }
	jade_engine* engine() { return m_engine; }
	/* newest columns as ARGB32 pixels coloured on the GPU (rows = getSpectrumSize(), row 0 = highest bin) */
	int fetchPixelColumns(uint32_t* pixels, int maxCols, int64_t* firstCol)
	{
		int n = 0;
		if (!m_engine || jade_fetch_columns(m_engine, pixels, nullptr, maxCols, &n, firstCol) != JADE_OK)
			return -1;
		m_newEntryCounter = 0;
		return n;
	}

private:
	static constexpr int kAllNew = int(100000000000LL % 4294967296LL); /* int(100000000000) on the reference toolchain */

	int setError(const char* msg)
	{
		m_error = msg ? msg : "error";
		return -1;
	}

	/* Spectrogram.cpp:213-238: every structural setter rebuilds buffers and resets the ring */
	void buildmem()
	{
		m_feedblocks = m_feed_percent == 100 ? 1 : m_feed_percent == 50 ? 2 : m_feed_percent == 25 ? 4 : 10;
		m_feed_samples = int(float(m_feed_percent) * 0.01 * m_fftsize + 0.5);
		m_memsize_blocks = int(m_memsize_s * m_fs / m_feed_samples + 0.5);
		m_freqsize = m_fftsize / 2 + 1;
		m_newEntryCounter = kAllNew;
		m_configured = false;
		if (!m_engine)
			return;
		jade_config c;
		jade_config_default(&c);
		c.sample_rate = m_fs;
		c.fft_size = int(m_fftsize);
		c.hop = m_feed_samples;
		c.frames_per_block = m_feedblocks;
		c.block_stride = int(m_fftsize);
		c.preroll = int(m_fftsize);
		c.emit_mode = JADE_EMIT_BLOCK;
		c.window = int(m_windowChoice);
		c.channels = int(m_channels);
		c.mix_mode = int(m_mode);
		c.row_map = JADE_ROWS_IDENTITY;
		c.flip_y = 1;
		c.pixel_format = JADE_PIX_ARGB32;
		c.memory_time_s = m_memsize_s;
		c.ring_columns = m_memsize_blocks > 0 ? m_memsize_blocks : 1;
		c.max_push = int(m_fftsize);
		if (jade_configure(m_engine, &c) != JADE_OK) {
			setError(jade_last_error(m_engine));
			return;
		}
		jade_set_pause(m_engine, m_PauseMode ? 1 : 0);
		m_configured = true;
	}

	jade_engine* m_engine = nullptr;
	bool m_configured = false;
	std::string m_error;
	float m_fs = 48000.0;
	size_t m_channels = 2;
	int m_feed_percent = 100;
	int m_feed_samples = 1024;
	int m_feedblocks = 1;
	float m_memsize_s = 1.0;
	int m_memsize_blocks = 0;
	size_t m_freqsize = 0;
	size_t m_fftsize = 1024;
	ChannelMixMode m_mode = ChannelMixMode::AbsMean;
	Windows m_windowChoice = Windows::Hann;
	int m_newEntryCounter = kAllNew;
	bool m_PauseMode = false;
	std::vector<float> m_scratch;
};
