/*
 * Stub of the JUCE surface that the reference's Spectrogram.cpp / PluginEditor.h / PluginProcessor.h touch.
 * TEST INFRASTRUCTURE ONLY (oracle/_ref build): lets the reference's REAL translation unit Spectrogram.cpp compile
 * where it lies under /root/reference, so that its own Spectrogram (framing, window, mix, dB, ring, getMem) and
 * SpectrogramComponent::timerCallback (pixel loops) run here and pin the restated oracle.
 * Widgets are inert value holders; Image is a real W x H ARGB32 array so the pixel loops have somewhere to write.
 * Nothing here is JUCE code: names and signatures only, written from the call sites in the reference.
 */
#pragma once
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <functional>
#include <map>
#include <memory>
#include <string>
#include <vector>

#define JUCE_DECLARE_NON_COPYABLE_WITH_LEAK_DETECTOR(cls) \
    cls(const cls&) = delete;                             \
    cls& operator=(const cls&) = delete;

namespace juce
{
using uint32 = std::uint32_t;
template <typename... T> void ignoreUnused(T&&...) {}

class String
{
public:
    String() {}
    String(const char* s) : m(s) {}
    String(const std::string& s) : m(s) {}
    explicit String(int v) : m(std::to_string(v)) {}
    explicit String(float v) : m(std::to_string(v)) {}
    explicit String(double v) : m(std::to_string(v)) {}
    String(double v, int) : m(std::to_string(v)) {}
    String& operator+=(const String& o) { m += o.m; return *this; }
    String& operator+=(const char* o) { m += o; return *this; }
    String operator+(const String& o) const { return String(m + o.m); }
    float getFloatValue() const { return float(std::atof(m.c_str())); }
    std::string m;
};

enum NotificationType { dontSendNotification = 0, sendNotification = 1 };

class Colour
{
public:
    Colour() : argb(0) {}
    explicit Colour(uint32 v) : argb(v) {}
    Colour darker(float = 0.4f) const { return *this; }
    uint32 getARGB() const { return argb; }
    uint32 argb;
};
namespace Colours
{
const Colour red(0xffff0000u);
const Colour white(0xffffffffu);
} // namespace Colours

struct Justification { enum Flags { centred = 36 }; Justification(int = 0) {} };

class Image
{
public:
    enum PixelFormat { UnknownFormat, RGB, ARGB, SingleChannel };
    Image() : w(0), h(0) {}
    Image(PixelFormat, int width, int height, bool) : w(width), h(height), px(size_t(width) * height, 0xff000000u) {}
    /* JUCE resamples; the reference only uses this to change the size before a full redraw (Spectrogram.cpp:598) */
    Image rescaled(int nw, int nh) const { return Image(RGB, nw, nh, true); }
    void setPixelAt(int x, int y, Colour c) { if (x >= 0 && y >= 0 && x < w && y < h) px[size_t(y) * w + x] = c.argb; }
    /* copies a rectangle inside the image: (dx,dy) <- (sx,sy) of size (width,height) */
    void moveImageSection(int dx, int dy, int sx, int sy, int width, int height)
    {
        std::vector<uint32> tmp(size_t(width > 0 ? width : 0) * (height > 0 ? height : 0));
        for (int y = 0; y < height; ++y)
            for (int x = 0; x < width; ++x) tmp[size_t(y) * width + x] = px[size_t(sy + y) * w + (sx + x)];
        for (int y = 0; y < height; ++y)
            for (int x = 0; x < width; ++x) px[size_t(dy + y) * w + (dx + x)] = tmp[size_t(y) * width + x];
    }
    class BitmapData
    {
    public:
        enum ReadWriteMode { readOnly, writeOnly, readWrite };
        BitmapData(Image& im, int, int, int, int, ReadWriteMode) : img(im) {}
        void setPixelColour(int x, int y, Colour c) const { img.setPixelAt(x, y, c); }
        Image& img;
    };
    int w, h;
    std::vector<uint32> px;
};

class LookAndFeel { public: Colour findColour(int) const { return Colour(0xff202020u); } };
struct ResizableWindow { enum ColourIds { backgroundColourId = 0x1005700 }; };

class MouseEvent { public: int getMouseDownX() const { return 0; } int getMouseDownY() const { return 0; } };
class MidiMessage { public: static String getMidiNoteName(int, bool, bool, int) { return String("C"); } };
class MidiBuffer {};
class MemoryBlock {};
class XmlElement {};
template <typename T> class AudioBuffer {};

/* records what paint() draws (text boxes and image blits) so that the axis / colourbar arithmetic of the reference's own
 * SpectrogramComponent::paint (Spectrogram.cpp:432-545) can be read back by oracle/ref_glue.cpp */
class Graphics
{
public:
    struct TextCall { std::string text; int x, y, w, h; };
    struct ImageCall { int dx, dy, dw, dh, sx, sy, sw, sh, iw, ih; std::vector<uint32> px; };
    void fillAll(Colour) {}
    void drawImage(const Image& im, int dx, int dy, int dw, int dh, int sx, int sy, int sw, int sh, bool = false)
    {
        images.push_back(ImageCall{dx, dy, dw, dh, sx, sy, sw, sh, im.w, im.h, im.w <= 2 ? im.px : std::vector<uint32>()});
    }
    void setFont(float) {}
    void drawText(const String& t, int x, int y, int w, int h, Justification, bool = true) { texts.push_back(TextCall{t.m, x, y, w, h}); }
    std::vector<TextCall> texts;
    std::vector<ImageCall> images;
};

class Component
{
public:
    virtual ~Component() {}
    virtual void paint(Graphics&) {}
    virtual void resized() {}
    void addAndMakeVisible(Component&) {}
    void addAndMakeVisible(Component*) {}
    LookAndFeel& getLookAndFeel() { static LookAndFeel l; return l; }
    int getWidth() const { return stubWidth; }
    int getHeight() const { return stubHeight; }
    int stubWidth = 800, stubHeight = 550; /* the plugin's minimum editor size (PlugInGUISettings.h:3-5) */
    void repaint() {}
    void setBounds(int, int, int, int) {}
    void setVisible(bool) {}
    void setColour(int, Colour) {}
};
class Timer
{
public:
    virtual ~Timer() {}
    virtual void timerCallback() = 0;
    void startTimer(int) {}
    void stopTimer() {}
};

class Label : public Component
{
public:
    enum ColourIds { backgroundColourId = 1, textColourId, outlineColourId };
    void setText(const String&, NotificationType) {}
    void setJustificationType(Justification) {}
};
class Slider : public Component
{
public:
    enum SliderStyle { LinearHorizontal, LinearVertical };
    void setSliderStyle(SliderStyle) {}
    double getValue() const { return value; }
    void setValue(double v, NotificationType = sendNotification) { value = v; }
    std::function<void()> onValueChange;
    double value = 0.0;
};
class TextButton : public Component
{
public:
    void setButtonText(const String&) {}
    void setToggleState(bool, NotificationType) {}
    std::function<void()> onClick;
};
class ComboBox : public Component
{
public:
    enum ColourIds { backgroundColourId = 1 };
    void addItem(const String&, int) {}
    void setSelectedItemIndex(int i, NotificationType = sendNotification) { sel = i; }
    int getSelectedItemIndex() const { return sel; }
    std::function<void()> onChange;
    int sel = 0;
};

template <typename T> class NormalisableRange { public: NormalisableRange(T, T) {} };
class AudioProcessorParameter { public: enum Category { genericParameter = 0 }; virtual ~AudioProcessorParameter() {} };
class RangedAudioParameter : public AudioProcessorParameter {};
class AudioParameterFloat : public RangedAudioParameter
{
public:
    AudioParameterFloat(const std::string&, const std::string&, NormalisableRange<float>, float, const std::string&,
                        AudioProcessorParameter::Category, std::function<String(float, int)>, std::function<float(const String&)>) {}
};

class AudioProcessorValueTreeState
{
public:
    std::atomic<float>* getRawParameterValue(const std::string& id) { return &vals[id]; }
    class SliderAttachment { public: SliderAttachment(AudioProcessorValueTreeState&, const std::string&, Slider&) {} };
    std::map<std::string, std::atomic<float>> vals;
};

class CriticalSection { public: void enter() const {} void exit() const {} };

class AudioProcessorEditor : public Component
{
public:
    virtual ~AudioProcessorEditor() {}
};
class AudioProcessor
{
public:
    struct BusesLayout {};
    virtual ~AudioProcessor() {}
    virtual void prepareToPlay(double, int) = 0;
    virtual void releaseResources() = 0;
    virtual bool isBusesLayoutSupported(const BusesLayout&) const { return true; }
    virtual void processBlock(AudioBuffer<float>&, MidiBuffer&) = 0;
    virtual void processBlock(AudioBuffer<double>&, MidiBuffer&) {}
    virtual AudioProcessorEditor* createEditor() = 0;
    virtual bool hasEditor() const = 0;
    virtual const String getName() const = 0;
    virtual bool acceptsMidi() const = 0;
    virtual bool producesMidi() const = 0;
    virtual bool isMidiEffect() const = 0;
    virtual double getTailLengthSeconds() const = 0;
    virtual int getNumPrograms() = 0;
    virtual int getCurrentProgram() = 0;
    virtual void setCurrentProgram(int) = 0;
    virtual const String getProgramName(int) = 0;
    virtual void changeProgramName(int, const String&) = 0;
    virtual void getStateInformation(MemoryBlock&) = 0;
    virtual void setStateInformation(const void*, int) = 0;
};
} // namespace juce
using namespace juce;
