/* acados_c/external_function_interface.h — include-path shim.  The reference's NMPCNavControl.h:10-17 includes this header but uses
 * no symbol from it; it exists so the wrapper compiles unchanged against the
 * B200-native solver. */
#ifndef NMPC_B200_SHIM_ACADOS_C_EXTERNAL_FUNCTION_INTERFACE_H
#define NMPC_B200_SHIM_ACADOS_C_EXTERNAL_FUNCTION_INTERFACE_H
#endif
