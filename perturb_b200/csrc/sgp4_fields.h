// Field schemas of the two structure-of-arrays layouts the batch keeps in HBM.
//
//  * RECORD layout  rec[RF_COUNT][n]       -- caller order. The numeric mirror of
//    sgp4::elsetrec (include/perturb/sgp4.hpp:89-134) as `sgp4init` leaves it; written by
//    kernel 1 (or uploaded pre-initialised), read back for parity / write-back to hosts.
//  * PROPAGATION layout  pf[PF_COUNT][n_pad] -- class-sorted order (near-Earth full drag,
//    near-Earth simplified, deep non-resonant, deep 24 h resonant, deep 12 h resonant; every
//    class padded to a multiple of 32 satellites). Holds only what `sgp4()` reads per
//    evaluation plus a few per-satellite products hoisted out of it (BC4 = bstar*cc4, ...).
//
// Plain C header: shared by the CUDA sources and the host-side C++.
#ifndef PERTURB_B200_SGP4_FIELDS_H
#define PERTURB_B200_SGP4_FIELDS_H

// X-macro lists keep enum, name table and count in one place.
// Order of the RECORD list == order of oracle/ref_driver.cpp:ref_dump_fields (a test
// convenience only; the product never depends on the oracle).
#define PB_RECORD_FIELDS(X)                                                                  \
    X(error) X(isimp) X(irez) X(method_d)                                                    \
    X(aycof) X(con41) X(cc1) X(cc4) X(cc5) X(d2) X(d3) X(d4) X(delmo) X(eta) X(argpdot)       \
    X(omgcof) X(sinmao) X(t2cof) X(t3cof) X(t4cof) X(t5cof) X(x1mth2) X(x7thm1) X(mdot)       \
    X(nodedot) X(xlcof) X(xmcof) X(nodecf)                                                   \
    X(d2201) X(d2211) X(d3210) X(d3222) X(d4410) X(d4422) X(d5220) X(d5232) X(d5421)          \
    X(d5433) X(dedt) X(del1) X(del2) X(del3) X(didt) X(dmdt) X(dnodt) X(domdt) X(e3) X(ee2)   \
    X(peo) X(pgho) X(pho) X(pinco) X(plo) X(se2) X(se3) X(sgh2) X(sgh3) X(sgh4) X(sh2)        \
    X(sh3) X(si2) X(si3) X(sl2) X(sl3) X(sl4) X(gsto) X(xfact) X(xgh2) X(xgh3) X(xgh4)        \
    X(xh2) X(xh3) X(xi2) X(xi3) X(xl2) X(xl3) X(xl4) X(xlamo) X(zmol) X(zmos) X(atime)        \
    X(xli) X(xni)                                                                            \
    X(a) X(altp) X(alta) X(jdsatepoch) X(jdsatepochF) X(nddot) X(ndot) X(bstar) X(inclo)      \
    X(nodeo) X(ecco) X(argpo) X(mo) X(no_kozai) X(no_unkozai)                                \
    X(am) X(em) X(im) X(Om) X(om) X(mm) X(nm)                                                \
    X(tumin) X(mus) X(radiusearthkm) X(xke) X(j2) X(j3) X(j4) X(j3oj2)

enum {
#define X(name) RF_##name,
    PB_RECORD_FIELDS(X)
#undef X
    RF_COUNT
};

// Propagation fields. Ranges: [0, PF_NEAR_END) near-Earth simplified drag,
// [0, PF_FULL_END) near-Earth full drag, [0, PF_COMMON_END) + [PF_FULL_END, PF_COUNT) deep.
#define PB_PROP_FIELDS(X)                                                                    \
    X(JD0) X(JDF0) X(MO) X(MDOT) X(ARGPO) X(ARGPDOT) X(NODEO) X(NODEDOT) X(NODECF) X(CC1)     \
    X(BC4) X(T2COF) X(NO) X(ECCO) X(INCLO) X(AOBASE) X(SINIO) X(COSIO)                       \
    /* near-Earth only */                                                                    \
    X(AYCOF) X(XLCOF) X(CON41) X(X1MTH2) X(X7THM1)                                           \
    /* near-Earth full drag only */                                                          \
    X(OMGCOF) X(ETA) X(XMCOF) X(DELMO) X(D2) X(D3) X(D4) X(BC5) X(SINMAO) X(T3COF) X(T4COF)   \
    X(T5COF)                                                                                 \
    /* deep space */                                                                         \
    X(DEDT) X(DIDT) X(DMDT) X(DNODT) X(DOMDT) X(GSTO)                                        \
    X(E3) X(EE2) X(SE2) X(SE3) X(SGH2) X(SGH3) X(SGH4) X(SH2) X(SH3) X(SI2) X(SI3) X(SL2)     \
    X(SL3) X(SL4) X(XGH2) X(XGH3) X(XGH4) X(XH2) X(XH3) X(XI2) X(XI3) X(XL2) X(XL3) X(XL4)    \
    X(ZMOL) X(ZMOS)                                                                          \
    /* deep space, resonant */                                                               \
    X(XFACT) X(XLAMO) X(DEL1) X(DEL2) X(DEL3)                                                \
    X(D2201) X(D2211) X(D3210) X(D3222) X(D4410) X(D4422) X(D5220) X(D5232) X(D5421) X(D5433)

enum {
#define X(name) PF_##name,
    PB_PROP_FIELDS(X)
#undef X
    PF_COUNT,
    PF_COMMON_END = PF_COSIO + 1,
    PF_NEAR_END = PF_X7THM1 + 1,
    PF_FULL_END = PF_T5COF + 1,
    PF_DEEP_NR_END = PF_ZMOS + 1,
    PF_DEEP_R1_END = PF_DEL3 + 1
};

// Orbit classes == sort keys == kernel specialisations.
enum {
    PB_CLS_NEAR_FULL = 0,  // method 'n', isimp 0
    PB_CLS_NEAR_SIMP = 1,  // method 'n', isimp 1   (perigee < 220 km, src/sgp4.cpp:1516)
    PB_CLS_DEEP_NR = 2,    // method 'd', irez 0
    PB_CLS_DEEP_R1 = 3,    // method 'd', irez 1    (24 h resonance, src/sgp4.cpp:773)
    PB_CLS_DEEP_R2 = 4,    // method 'd', irez 2    (12 h resonance, src/sgp4.cpp:775)
    PB_CLS_COUNT = 5,
    PB_CLS_RUNTIME = -1    // flags taken from the record at run time (kernel 1's sgp4(t=0))
};

#define PB_TILE_SATS 32  // satellites per tile: class segments are padded to this

#endif
