/* Stub of the author's TGM PresetHandler.h (absent).  TEST INFRASTRUCTURE ONLY. */
#pragma once
#include <juce_audio_processors/juce_audio_processors.h>
class PresetHandler {};
class PresetComponent : public juce::Component {};
