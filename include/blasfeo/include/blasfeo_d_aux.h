/* blasfeo/include/blasfeo_d_aux.h — include-path shim.  The reference's NMPCNavControl.h:10-17 includes this header but uses
 * no symbol from it; it exists so the wrapper compiles unchanged against the
 * B200-native solver. */
#ifndef NMPC_B200_SHIM_BLASFEO_INCLUDE_BLASFEO_D_AUX_H
#define NMPC_B200_SHIM_BLASFEO_INCLUDE_BLASFEO_D_AUX_H
#endif
