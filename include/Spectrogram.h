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
 * With JUCE present the header also carries what the reference's Spectrogram.h declares for the plugin shell and that
 * PluginProcessor.{h,cpp} use: the four parameter descriptors (Spectrogram.h:22-58), `SpectrogramParameter`
 * (Spectrogram.h:61-76, Spectrogram.cpp:793-832) and `Spectrogram::prepareParameter` (Spectrogram.h:113,
 * Spectrogram.cpp:25-35, called at PluginProcessor.cpp:28).  `SpectrogramComponent` (Spectrogram.h:171-236, GUI) is NOT
 * here: INTEGRATION.md section 1 says where it goes.
 *
 * Threading: one audio thread (processSynchronBlock) and one GUI thread (getMem, setters), as in the plugin.
 * m_newEntryCounter is atomic; the structural setters (buildmem) hold m_protect, processSynchronBlock only try-locks it
 * and drops the block while a rebuild is in progress (the rebuild clears all state anyway) -- the audio thread never
 * waits for the GUI thread.
 *
 * There is no CPU implementation behind this class: without a CUDA device every processing call returns -1 and
 * lastError() says why (the reference's int-status convention; no exceptions cross the audio thread).
 */
#pragma once

#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <memory>
#include <mutex>
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
#if __has_include("PlugInGUISettings.h")
#include "PlugInGUISettings.h" /* g_minColorVal / g_maxColorVal (PlugInGUISettings.h:37-38) */
#define JADE_MIN_COLOR_VAL float(g_minColorVal)
#define JADE_MAX_COLOR_VAL float(g_maxColorVal)
#else
#define JADE_MIN_COLOR_VAL (-50.0f)
#define JADE_MAX_COLOR_VAL (50.0f)
#endif

/* Host-automatable display parameters of the plugin (Spectrogram.h:22-58): the frequency limits are stored as log(Hz), the
 * colour limits in dB.  Same object names and fields as the reference, because the component code keeps using them. */
struct JadeDisplayParamDesc
{
	std::string ID, name, unitName;
	float minValue, maxValue, defaultValue;
};
inline const JadeDisplayParamDesc paramDisplayMinFreq{"MinFreq", "MinFreq", "Hz", std::log(1.f), std::log(10000.f), std::log(1.f)};
inline const JadeDisplayParamDesc paramDisplayMaxFreq{"MaxFreq", "MaxFreq", "Hz", std::log(500.f), std::log(20000.f), std::log(20000.f)};
inline const JadeDisplayParamDesc paramDisplayMinColor{"MinColor", "MinColor", "", JADE_MIN_COLOR_VAL, JADE_MAX_COLOR_VAL, JADE_MIN_COLOR_VAL};
inline const JadeDisplayParamDesc paramDisplayMaxColor{"MaxColor", "MaxColor", "", JADE_MIN_COLOR_VAL, JADE_MAX_COLOR_VAL, JADE_MAX_COLOR_VAL};

/* Spectrogram.h:61-76 / Spectrogram.cpp:793-832 */
class SpectrogramParameter
{
public:
	SpectrogramParameter() {}
	/* appends the four AudioParameterFloat objects: frequencies shown as Hz with one decimal (stored value is log Hz),
	 * colour limits as whole dB */
	int addParameter(std::vector<std::unique_ptr<RangedAudioParameter>>& paramVector)
	{
		const JadeDisplayParamDesc* descs[4] = {&paramDisplayMinFreq, &paramDisplayMaxFreq, &paramDisplayMinColor, &paramDisplayMaxColor};
		for (int i = 0; i < 4; ++i) {
			const JadeDisplayParamDesc& d = *descs[i];
			const bool isFreq = i < 2;
			paramVector.push_back(std::make_unique<AudioParameterFloat>(
				d.ID, d.name, NormalisableRange<float>(d.minValue, d.maxValue), d.defaultValue, d.unitName,
				AudioProcessorParameter::genericParameter,
				[isFreq](float value, int maxLen) {
					return isFreq ? String(0.1 * int(std::exp(value) * 10 + 0.5), maxLen) : String(1.0 * int(value + 0.5), maxLen);
				},
				[](const String& text) { return text.getFloatValue(); }));
		}
		return 0;
	}

	std::atomic<float>* m_DisplayMinFreq = nullptr;
	float m_DisplayMinFreqOld = 0.f;
	std::atomic<float>* m_DisplayMaxFreq = nullptr;
	float m_DisplayMaxFreqOld = 0.f;
	std::atomic<float>* m_DisplayMinColor = nullptr;
	float m_DisplayMinColorOld = 0.f;
	std::atomic<float>* m_DisplayMaxColor = nullptr;
	float m_DisplayMaxColorOld = 0.f;
};
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
		std::lock_guard<std::mutex> lk(m_accLock);
		m_NrOfChannels = channels > 0 ? channels : 1;
		resetAccumulator();
	}
	void setDesiredBlockSizeSamples(int n)
	{
		std::lock_guard<std::mutex> lk(m_accLock);
		m_desired = n > 0 ? n : 1;
		resetAccumulator();
	}
	virtual int processSynchronBlock(std::vector<std::vector<float>>&, juce::MidiBuffer&) = 0;

	/* planar channel pointers, any numSamples */
	int processBlock(const float* const* channelData, int numChannels, int numSamples, juce::MidiBuffer& midi)
	{
		/* the audio thread never waits: while the GUI thread resizes the accumulator the host block is dropped */
		std::unique_lock<std::mutex> lk(m_accLock, std::try_to_lock);
		if (!lk.owns_lock())
			return 0;
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
	std::mutex m_accLock; /* accumulator vs preparetoProcess / setDesiredBlockSizeSamples from another thread */
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

#if defined(JADE_HAVE_TGM_JUCE)
	/* Spectrogram.cpp:25-35 (call site PluginProcessor.cpp:28): raw parameter pointers of the value tree state */
	void prepareParameter(std::unique_ptr<AudioProcessorValueTreeState>& vts)
	{
		const JadeDisplayParamDesc* descs[4] = {&paramDisplayMinFreq, &paramDisplayMaxFreq, &paramDisplayMinColor, &paramDisplayMaxColor};
		std::atomic<float>** raw[4] = {&m_SpecParameter.m_DisplayMinFreq, &m_SpecParameter.m_DisplayMaxFreq,
									   &m_SpecParameter.m_DisplayMinColor, &m_SpecParameter.m_DisplayMaxColor};
		float* old[4] = {&m_SpecParameter.m_DisplayMinFreqOld, &m_SpecParameter.m_DisplayMaxFreqOld,
						 &m_SpecParameter.m_DisplayMinColorOld, &m_SpecParameter.m_DisplayMaxColorOld};
		for (int i = 0; i < 4; ++i) {
			*raw[i] = vts->getRawParameterValue(descs[i]->ID);
			*old[i] = descs[i]->defaultValue;
		}
	}
#endif

	/* Spectrogram.cpp:37-135.  data[channel][fftsize]; returns 0 (the reference always does) or -1 on error. */
	virtual int processSynchronBlock(std::vector<std::vector<float>>& data, juce::MidiBuffer& midiMessages)
	{
		(void)midiMessages;
		std::unique_lock<std::mutex> lk(m_protect, std::try_to_lock);
		if (!lk.owns_lock())
			return 0; /* a structural setter is rebuilding the engine: this block is dropped, the rebuild clears all state */
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
			m_newEntryCounter.fetch_add(m_feedblocks, std::memory_order_release);
		return 0;
	}

	// setter (Spectrogram.cpp:148-211)
	void setSamplerate(float samplerate)
	{
		std::lock_guard<std::mutex> lk(m_protect);
		m_fs = samplerate;
		buildmem();
	}
	void setchannels(size_t newchannels)
	{
		std::lock_guard<std::mutex> lk(m_protect);
		m_channels = newchannels;
		buildmem();
	}
	void setFFTSize(size_t newFFTSize)
	{
		std::lock_guard<std::mutex> lk(m_protect); /* the reference's m_protect.enter(), Spectrogram.cpp:162 */
		m_fftsize = newFFTSize;
		setDesiredBlockSizeSamples(int(m_fftsize));
		buildmem();
		m_newEntryCounter = kAllNew;
	}
	void setclosestFFTSize_ms(float fftsize_ms)
	{
		std::lock_guard<std::mutex> lk(m_protect);
		m_fftsize = getnextpowerof2(fftsize_ms);
		setDesiredBlockSizeSamples(int(m_fftsize));
		buildmem();
	}
	void setmemoryTime_s(float memsize_s)
	{
		std::lock_guard<std::mutex> lk(m_protect);
		m_memsize_s = memsize_s;
		buildmem();
	}
	void setfeed_percent(FeedPercentage feed)
	{
		std::lock_guard<std::mutex> lk(m_protect);
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
		std::lock_guard<std::mutex> lk(m_protect);
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
		/* The counter is read and cleared in ONE atomic step (the reference reads, uses and zeroes a plain int here,
		 * Spectrogram.cpp:295-331).  WHICH columns are new is decided by the engine's own fetch cursor, so a block the audio
		 * thread pushes between two statements of this function is neither lost nor delivered twice: it is either part of
		 * this call or of the next one (a counter increment that arrives after its column was delivered only makes the next call
		 * look, and find nothing). */
		const int counted = m_newEntryCounter.exchange(0, std::memory_order_acq_rel);
		int n = 0;
		int64_t first = 0;
		if (counted >= W) { /* :299-304 -- everything is new: the whole ring */
			if (jade_fetch_columns(m_engine, nullptr, nullptr, 0, &n, &first) != JADE_OK) /* first mark everything as seen ... */
				return setError(jade_last_error(m_engine));
			m_scratch.resize(size_t(W) * B);
			if (jade_read_ring_db(m_engine, m_scratch.data()) != JADE_OK)                 /* ... then read at least that much */
				return setError(jade_last_error(m_engine));
			for (int kk = 0; kk < W; ++kk)
				std::copy(m_scratch.begin() + size_t(kk) * B, m_scratch.begin() + size_t(kk + 1) * B, mem[size_t(kk)].begin());
			pos = int(first % W);
			return counted;
		}
		/* :305-320 -- the newest columns into their ring slots */
		m_scratch.resize(size_t(W) * B);
		if (jade_fetch_columns(m_engine, nullptr, m_scratch.data(), W, &n, &first) != JADE_OK)
			return setError(jade_last_error(m_engine));
		for (int i = 0; i < n; ++i) {
			const size_t slot = size_t((first + i) % W);
			std::copy(m_scratch.begin() + size_t(i) * B, m_scratch.begin() + size_t(i + 1) * B, mem[slot].begin());
		}
		pos = int((first + n) % W);
		return n;
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
		m_newEntryCounter.store(0, std::memory_order_release);
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
	std::atomic<int> m_newEntryCounter{kAllNew};
	std::atomic<bool> m_PauseMode{false};
	std::mutex m_protect; /* structural setters vs processSynchronBlock (the reference's CriticalSection, Spectrogram.h:133) */
#if defined(JADE_HAVE_TGM_JUCE)
	SpectrogramParameter m_SpecParameter; /* Spectrogram.h:165 */
#endif
	std::vector<float> m_scratch;
};
