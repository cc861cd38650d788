/* dimensions.h - index macros and Cartesian ordering that callers of the reference include
 * (reference src/dimensions.h:4-40; example/ex1.c:7 includes it, so it is part of the de-facto API).
 * Macro names and values are the reference's; fully parenthesised here. */
#ifndef LIBECP_DIMENSIONS_H
#define LIBECP_DIMENSIONS_H 1

#define L_QN(l) (l)
#define M_INDEX_MIN(l) (-(l))
#define M_INDEX_MAX(l) (+(l))
#define M_QN(l,mIndex) (M_INDEX_MIN(l)+(mIndex))
#define M_INDEX(l,m) ((m)-M_INDEX_MIN(l))

/* l:        0  1  2  3  4
   L_DIM:    1  4  9 16 25
   M_DIM:    1  3  5  7  9 */
#define L_DIM(l) (((l)+1)*((l)+1))
#define M_DIM(l) (2*(l)+1)
#define LM_INDEX(l,m) ((l)*(l)+(m))

/* l:        0  1  2  3  4
   IJK_DIM:  1  3  6 10 15
   C_DIM:    1  4 10 20 35 */
#define C_DIM(l) (((l)+1)*((l)+2)*((l)+3)/6)
#define C_INDEX(l,c) (C_DIM((l)-1)+(c))
#define IJK_DIM(l) (((l)+1)*((l)+2)/2)
#define CIJK_INDEX(l,c) ((C_DIM((l)-1)+(c))*3)

#ifdef __cplusplus
extern "C" {
#endif
/* exponents nx,ny,nz of the Cartesian shells up to am in libint order (p: x,y,z; d: xx,xy,xz,yy,yz,zz; ...),
 * 3*C_DIM(am) ints, caller frees (reference src/dimensions.c:17-38) */
int * cartesianShellOrder(const int am);
/* inverse table [(am+1)^3]: (i,j,k) -> C_INDEX, caller frees (reference src/dimensions.c:41-57) */
int * cartesianShellOrderIndex(const int am, int *ijk);
#ifdef __cplusplus
}
#endif
#endif
