// plugin_compile_test.cpp -- compile-only check that the plugin shell's use of the Spectrogram header builds against
// include/Spectrogram.h when JUCE and the TGM headers are on the include path (here: the inert stubs in oracle/shim).
// It repeats, statement for statement, what the reference's PluginProcessor does with the class: the members declared at
// PluginProcessor.h:56-64, the constructor statements of PluginProcessor.cpp:21-28, prepareToPlay (:102-114) and
// processBlock (:145-150).  TEST INFRASTRUCTURE ONLY.
#include <juce_audio_processors/juce_audio_processors.h>

#include "PresetHandler.h"
#include "Spectrogram.h"

#if !defined(JADE_HAVE_TGM_JUCE)
#error "the JUCE / TGM branch of include/Spectrogram.h was not selected"
#endif

struct PluginShell {
    std::unique_ptr<AudioProcessorValueTreeState> m_parameterVTS;
    std::vector<std::unique_ptr<RangedAudioParameter>> m_paramVector;
    PresetHandler m_presets;
    Spectrogram m_spectrogram;
    SpectrogramParameter m_specParameter;
    int m_fftsize = 2048;

    PluginShell()
    {
        m_specParameter.addParameter(m_paramVector);                          // PluginProcessor.cpp:21
        m_parameterVTS = std::make_unique<AudioProcessorValueTreeState>();    // :22-23 (the stub has no layout argument)
        m_spectrogram.prepareParameter(m_parameterVTS);                       // :28
    }
    void prepareToPlay(double sampleRate, int samplesPerBlock)                // :102-114
    {
        m_spectrogram.preparetoProcess(2, samplesPerBlock);
        m_spectrogram.setSamplerate(float(sampleRate));
        m_spectrogram.setmemoryTime_s(10.0);
        m_spectrogram.setFFTSize(size_t(m_fftsize));
        m_spectrogram.setfeed_percent(Spectrogram::FeedPercentage::perc50);
    }
    void processBlock(juce::AudioBuffer<float>& buffer, juce::MidiBuffer& midi) { m_spectrogram.processBlock(buffer, midi); } // :148
};

int plugin_shell_parameters() // the descriptors keep the reference's names and values (Spectrogram.h:22-58)
{
    PluginShell p;
    return int(p.m_paramVector.size()) + int(paramDisplayMinFreq.ID.size()) + int(paramDisplayMaxColor.defaultValue);
}
