/* Stub of the author's TGM SynchronBlockProcessor (absent from /root/reference): only the members Spectrogram.cpp
 * and PluginProcessor.cpp
 * use (call sites Spectrogram.cpp:164,180; PluginProcessor.cpp:108,148).  TEST INFRASTRUCTURE ONLY. */
#pragma once
#include <juce_audio_processors/juce_audio_processors.h>
class SynchronBlockProcessor
{
public:
    SynchronBlockProcessor() {}
    virtual ~SynchronBlockProcessor() {}
    void preparetoProcess(int, int) {}
    void setDesiredBlockSizeSamples(int n) { m_desired = n; }
    int processBlock(juce::AudioBuffer<float>&, juce::MidiBuffer&) { return 0; } /* PluginProcessor.cpp:148; inert here */
    virtual int processSynchronBlock(std::vector<std::vector<float>>&, juce::MidiBuffer&) = 0;
    int m_desired = 0;
};
