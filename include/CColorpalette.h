/*
 * CColorpalette.h -- drop-in replacement for the reference's CColorPalette (CColorpalette.h:6-63, CColorpalette.cpp).
 *
 * Same public surface (enum, constructors, setValueRange / setNrOfColors / setColorSceme / setInvertStatus,
 * getRGBColor, getValue), same table contents and lookup arithmetic -- verified bit-exact against the reference's own
 * CColorpalette.cpp compiled in place (tests/test_palette.py).  The colour tables come from jade_palette_build in
 * libjade_gpu.so; the per-pixel lookup of whole columns runs on the GPU inside the STFT kernels (the table and the
 * value range are mirrored into an engine with bindEngine()), while the scalar getRGBColor below serves the few
 * host-side calls the GUI makes (colour bar, Spectrogram.cpp:511-521).
 */
#pragma once

#include <cstdint>
#include <vector>

#include "jade_gpu.h"

class CColorPalette
{
public:
	enum
	{
		kMono = 0,
		kBW,
		kHot,
		kRainbow,
		kViridis,
		kPlasma,
		kJade
	};

	/* The reference declares CColorPalette(int) and CColorPalette(int, int = kRainbow), which makes every one-argument
	 * call ambiguous (CColorpalette.h:21-22); only the zero- and two-argument forms are callable there. */
	CColorPalette() : m_NrOfColors(2), m_ColorScheme(kMono) { init(); }
	CColorPalette(int NrOfColors, int ColorScheme) : m_NrOfColors(NrOfColors), m_ColorScheme(ColorScheme) { init(); }
	~CColorPalette() {}

	// setter
	void setValueRange(float Min, float Max) /* CColorpalette.cpp:39-54 */
	{
		m_Min = Max >= Min ? Min : Max;
		m_Max = Max >= Min ? Max : Min;
		if (m_Max == m_Min)
			m_Min = float(0.99 * m_Max);
		m_AccessMult = float(m_NrOfColors) / (m_Max - m_Min);
		if (m_engine)
			jade_set_value_range(m_engine, m_Min, m_Max);
	}
	void setNrOfColors(int NrOfColors) /* :55-61 */
	{
		m_NrOfColors = NrOfColors;
		m_AccessMult = float(m_NrOfColors) / (m_Max - m_Min);
		AllocateColors();
	}
	void setColorSceme(int ColorScheme) /* :62-66 */
	{
		m_ColorScheme = ColorScheme;
		ComputeColors();
	}
	void setInvertStatus(bool status) { m_InvertScheme = status; } /* no recompute, like the reference */

	// Access (CColorpalette.h:32-47)
	inline int getRGBColor(float value)
	{
		float v = value >= m_Max ? m_Max * 0.9999f : value;
		v = v < m_Min ? m_Min : v;
		const int index = int((v - m_Min) * m_AccessMult);
		return m_Color[index < m_NrOfColors ? index : m_NrOfColors - 1];
	}
	float getValue(int iColor) /* :82-94 */
	{
		for (int kk = 0; kk < m_NrOfColors; kk++)
			if (m_Color[kk] == iColor)
				return float(kk) / m_AccessMult + m_Min;
		return 100000000000000000000000000000.f;
	}

	// ---- extension: mirror table + range into a GPU engine (columns are coloured inside the CUDA kernels) ----
	void bindEngine(jade_engine* e)
	{
		m_engine = e;
		push();
	}
	const std::vector<int>& table() const { return m_Color; }
	float getMin() const { return m_Min; }
	float getMax() const { return m_Max; }
	int getNrOfColors() const { return m_NrOfColors; }

protected:
	void init()
	{
		m_Min = 0.f;
		m_Max = 1.f;
		m_AccessMult = float(m_NrOfColors) / (m_Max - m_Min);
		AllocateColors();
	}
	void ComputeColors(void)
	{
		/* in-place rebuild: entries the scheme does not write keep their previous content, as in the reference */
		static_assert(sizeof(int) == sizeof(int32_t), "int must be 32 bit");
		if (!m_Color.empty())
			jade_palette_build(m_ColorScheme, m_NrOfColors, m_InvertScheme ? 1 : 0, reinterpret_cast<int32_t*>(m_Color.data()));
		push();
	}
	void AllocateColors(void)
	{
		m_Color.resize(m_NrOfColors);
		ComputeColors();
	}
	void push()
	{
		if (!m_engine || m_Color.empty())
			return;
		jade_set_palette(m_engine, reinterpret_cast<const int32_t*>(m_Color.data()), m_NrOfColors);
		jade_set_value_range(m_engine, m_Min, m_Max);
	}
	std::vector<int> m_Color;
	int m_NrOfColors;
	float m_Max;
	float m_Min;
	float m_AccessMult;
	int m_ColorScheme;
	int m_InvertScheme = 0;
	jade_engine* m_engine = nullptr;
};
