/* Stub of the author's TGM JadeLookAndFeel.h (absent).  TEST INFRASTRUCTURE ONLY. */
#pragma once
#include <juce_audio_processors/juce_audio_processors.h>
class JadeLookAndFeel : public juce::LookAndFeel {};
const juce::Colour JadeTeal(0xff0d9ba2u);
