/* unused_tables.cpp -- TEST INFRASTRUCTURE ONLY.
 * RGBToSpectrumTable::init (reference src/render/color.cpp:60-112) copies the coefficient tables of
 * all four colour spaces, but the wavefront path of this build only ever uses sRGB (the C ABI rejects
 * any other MaterialData colour space).  The three unused 9.4 MB tables (src/data/rgbspectrum_
 * {aces,dci_p3,rec2020}.cpp) are therefore NOT compiled into oracle/_ref: these zero-filled
 * definitions (in .bss, no file size) satisfy the linker so the reference's init() runs unmodified.
 * The sRGB table IS the reference's own src/data/rgbspectrum_srgb.cpp. */
namespace krr {
int	  DCI_P3ToSpectrumTable_Res = 64;
float DCI_P3ToSpectrumTable_Scale[64];
float DCI_P3ToSpectrumTable_Data[3][64][64][64][3];
int	  REC2020ToSpectrumTable_Res = 64;
float REC2020ToSpectrumTable_Scale[64];
float REC2020ToSpectrumTable_Data[3][64][64][64][3];
int	  ACES2065_1ToSpectrumTable_Res = 64;
float ACES2065_1ToSpectrumTable_Scale[64];
float ACES2065_1ToSpectrumTable_Data[3][64][64][64][3];
} // namespace krr
