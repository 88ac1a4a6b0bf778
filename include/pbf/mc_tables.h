/*
 * mc_tables.h — marching-cubes triangle table (Paul Bourke's public-domain "Polygonising a scalar
 * field" table, the same data the reference carries as TriTable in src/mc_constants.h:24-155),
 * re-packed for this code base: ONE 64-bit word per cube configuration, nibble t = t-th edge index
 * (0..11) of the triangle list, 0xF = end of list.  The reference's EdgeTable (mc_constants.h:4-20)
 * and NumVertsTable (:157-164) are DERIVED from it here instead of stored:
 *   edge mask  = OR of (1 << e) over the listed edges  (tests/test_mc_tables.py proves this equals the
 *                geometric definition "edge endpoints lie on different sides")
 *   vertex cnt = number of non-0xF nibbles.
 * Shared by the CUDA kernels (csrc/mc.cu) and by the CPU oracle (oracle/pbf_oracle.c).
 */
#ifndef PBF_MC_TABLES_H
#define PBF_MC_TABLES_H
#include <stdint.h>

#define PBF_MC_TRI_WORDS_INIT { \
  0xffffffffffffffffull, 0xfffffffffffff380ull, 0xfffffffffffff910ull, 0xffffffffff189381ull, \
  0xfffffffffffffa21ull, 0xffffffffffa21380ull, 0xffffffffff920a29ull, 0xfffffff89a8a2382ull, \
  0xfffffffffffff2b3ull, 0xffffffffff0b82b0ull, 0xffffffffffb32091ull, 0xfffffffb89b912b1ull, \
  0xffffffffff3ab1a3ull, 0xfffffffab8a801a0ull, 0xfffffff9ab9b3093ull, 0xffffffffffb8aa89ull, \
  0xfffffffffffff874ull, 0xffffffffff437034ull, 0xffffffffff748910ull, 0xfffffff137174914ull, \
  0xffffffffff748a21ull, 0xfffffffa21403743ull, 0xfffffff748209a29ull, 0xffff4973727929a2ull, \
  0xffffffffff2b3748ull, 0xfffffff40242b74bull, 0xfffffffb32748109ull, 0xffff1292b9b49b74ull, \
  0xfffffff487ab31a3ull, 0xffff4b7401b41ab1ull, 0xffff30bab9b09874ull, 0xfffffffab99b4b74ull, \
  0xfffffffffffff459ull, 0xffffffffff380459ull, 0xffffffffff051450ull, 0xfffffff513538458ull, \
  0xffffffffff459a21ull, 0xfffffff594a21803ull, 0xfffffff204245a25ull, 0xffff8434535235a2ull, \
  0xffffffffffb32459ull, 0xfffffff594b802b0ull, 0xfffffffb32510450ull, 0xffff584b82852512ull, \
  0xfffffff45931ab3aull, 0xffffab81a8180594ull, 0xffff30bab5b05045ull, 0xfffffffb8aa85845ull, \
  0xffffffffff975879ull, 0xfffffff375359039ull, 0xfffffff751710870ull, 0xffffffffff753351ull, \
  0xfffffff21a759879ull, 0xffff37503505921aull, 0xffff25a758528208ull, 0xfffffff7533525a2ull, \
  0xfffffff2b3987597ull, 0xffffb72029279759ull, 0xffff751871810b32ull, 0xfffffff51771b12bull, \
  0xffffb3a31a758859ull, 0xf0aba010b7905075ull, 0xf07570805a30b0abull, 0xffffffffff5b75abull, \
  0xfffffffffffff56aull, 0xffffffffff6a5380ull, 0xffffffffff6a5109ull, 0xfffffff6a5891381ull, \
  0xffffffffff162561ull, 0xfffffff803621561ull, 0xfffffff620609569ull, 0xffff823625285895ull, \
  0xffffffffff56ab32ull, 0xfffffff56a02b80bull, 0xfffffff6a5b32910ull, 0xffffb892b92916a5ull, \
  0xfffffff315356b36ull, 0xffff6b51505b0b80ull, 0xffff9505606306b3ull, 0xfffffff89bb96956ull, \
  0xffffffffff8746a5ull, 0xfffffffa56374034ull, 0xfffffff7486a5091ull, 0xffff49737179156aull, \
  0xfffffff874156216ull, 0xffff743403625521ull, 0xffff620560509748ull, 0xf962695923497937ull, \
  0xfffffff56a4872b3ull, 0xffffb720242746a5ull, 0xffff6a5b32874910ull, 0xf6a54b7b492b9129ull, \
  0xffff6b51535b3748ull, 0xfb404b7b016b5b15ull, 0xf74836b630560950ull, 0xffff9b7974b96956ull, \
  0xffffffffffa4694aull, 0xfffffff380a946a4ull, 0xfffffff04606a10aull, 0xffffa16468618138ull, \
  0xfffffff462421941ull, 0xffff462942921803ull, 0xffffffffff624420ull, 0xfffffff624428238ull, \
  0xfffffff32b46a94aull, 0xffff6a4a94b82280ull, 0xffffa164606102b3ull, 0xf1b8b12184a16146ull, \
  0xffff36b319639469ull, 0xf14641916b0181b8ull, 0xfffffff4600636b3ull, 0xffffffffff86b846ull, \
  0xfffffffa98a876a7ull, 0xffffa76a907a0370ull, 0xffff0818717a176aull, 0xfffffff37117a76aull, \
  0xffff768981861621ull, 0xf937390976192962ull, 0xfffffff206607087ull, 0xffffffffff276237ull, \
  0xffff76898a86ab32ull, 0xf7a9a76790b72702ull, 0xfb32a767a1871081ull, 0xffff17616a71b12bull, \
  0xf63136b619768698ull, 0xffffffffff76b190ull, 0xffff06b0b3607087ull, 0xfffffffffffff6b7ull, \
  0xfffffffffffffb67ull, 0xffffffffff67b803ull, 0xffffffffff67b910ull, 0xfffffff67b138918ull, \
  0xffffffffff7b621aull, 0xfffffff7b6803a21ull, 0xfffffff7b69a2092ull, 0xffff89a38a3a27b6ull, \
  0xffffffffff726327ull, 0xfffffff026067807ull, 0xfffffff910732672ull, 0xffff678891681261ull, \
  0xfffffff73171a67aull, 0xffff801781a7167aull, 0xffff7a69a0a70730ull, 0xfffffff9a88a7a67ull, \
  0xffffffffff68b486ull, 0xfffffff640603b63ull, 0xfffffff109648b68ull, 0xffff63b139369649ull, \
  0xfffffff1a28b6486ull, 0xffff640b60b03a21ull, 0xffff9a2920b648b4ull, 0xf36463b34923a39aull, \
  0xfffffff264248328ull, 0xffffffffff264240ull, 0xffff834642432091ull, 0xfffffff642241491ull, \
  0xffff1a6648168318ull, 0xfffffff40660a01aull, 0xf39a9303a6834364ull, 0xffffffffff4a649aull, \
  0xffffffffffb67594ull, 0xfffffff67b594380ull, 0xfffffffb67045105ull, 0xffff51345343867bull, \
  0xfffffffb6721a459ull, 0xffff594380a217b6ull, 0xffff204a24a45b67ull, 0xf67b25a523453843ull, \
  0xfffffff945267327ull, 0xffff786260680459ull, 0xffff045051673263ull, 0xf851584812786826ull, \
  0xffff73167161a459ull, 0xf459078701671a61ull, 0xfa737a6a305a4a04ull, 0xffffa84a458a7a67ull, \
  0xfffffff98b9b6596ull, 0xffff590650360b63ull, 0xffffb65510b508b0ull, 0xfffffff1355363b6ull, \
  0xffff65b8b9b59a21ull, 0xfa21965690b603b0ull, 0xf52025a50865b58bull, 0xffff35a3a25363b6ull, \
  0xffff283265825985ull, 0xfffffff260069659ull, 0xf826283865081851ull, 0xffffffffff612651ull, \
  0xf698965683a61631ull, 0xffff06505960a01aull, 0xffffffffffa65830ull, 0xfffffffffffff65aull, \
  0xffffffffffb57a5bull, 0xfffffff03857ba5bull, 0xfffffff091ba57b5ull, 0xffff1381897ba57aull, \
  0xfffffff15717b21bull, 0xffffb27571721380ull, 0xffff7b2209729579ull, 0xf289823295b27257ull, \
  0xfffffff573532a52ull, 0xffff52a578258028ull, 0xffff2a37353a5109ull, 0xf25752a278129289ull, \
  0xffffffffff573531ull, 0xfffffff571170780ull, 0xfffffff735539309ull, 0xffffffffff795789ull, \
  0xfffffff8ba8a5485ull, 0xffff03bba50b5405ull, 0xffff54aba8a48910ull, 0xf41314943b54a4baull, \
  0xffff8548b2582152ull, 0xfb151b2b543b0b40ull, 0xf58b8545b2950520ull, 0xffffffffff3b2549ull, \
  0xffff483543253a52ull, 0xfffffff0244252a5ull, 0xf910854583a532a3ull, 0xffff2492914252a5ull, \
  0xfffffff153358548ull, 0xffffffffff501540ull, 0xffff530509358548ull, 0xfffffffffffff549ull, \
  0xfffffffba9b947b4ull, 0xffffba97b9794380ull, 0xffffb470414b1ba1ull, 0xf4bab474a1843413ull, \
  0xffff219b294b97b4ull, 0xf3801b2b197b9479ull, 0xfffffff04224b47bull, 0xffff42343824b47bull, \
  0xffff947732972a92ull, 0xf70207872a4797a9ull, 0xfa040a1a472a3a73ull, 0xffffffffff4782a1ull, \
  0xfffffff317714194ull, 0xffff178180714194ull, 0xffffffffff347304ull, 0xfffffffffffff784ull, \
  0xffffffffff8ba8a9ull, 0xfffffffa9bb93903ull, 0xfffffffba88a0a10ull, 0xffffffffffa3ba13ull, \
  0xfffffff8b99b1b21ull, 0xffff9b2921b93903ull, 0xffffffffffb08b20ull, 0xfffffffffffffb23ull, \
  0xfffffff98aa82832ull, 0xffffffffff2902a9ull, 0xffff8a1810a82832ull, 0xfffffffffffff2a1ull, \
  0xffffffffff819831ull, 0xfffffffffffff190ull, 0xfffffffffffff830ull, 0xffffffffffffffffull \
}

/* The 12 cube edges as corner pairs, in the order of ompsph.hpp:443-454. */
#define PBF_MC_EDGE_CORNERS_INIT {{0,1},{1,2},{2,3},{3,0},{4,5},{5,6},{6,7},{7,4},{0,4},{1,5},{2,6},{3,7}}
/* The 8 cube corners as lattice offsets, in the order of ompsph.hpp:361-363. */
#define PBF_MC_CORNER_OFFSETS_INIT {{0,0,0},{1,0,0},{1,1,0},{0,1,0},{0,0,1},{1,0,1},{1,1,1},{0,1,1}}

#if defined(__CUDACC__)
#define PBF_MC_HD __host__ __device__ __forceinline__
#else
#define PBF_MC_HD static inline
#endif

/* number of triangle vertices listed in a packed row */
PBF_MC_HD uint32_t pbf_mc_num_verts(uint64_t row) {
  uint32_t n = 0;
  while (n < 16u && ((row >> (4u * n)) & 0xFu) != 0xFu) ++n;
  return n;
}
/* 12-bit mask of the edges a packed row uses (== the reference's EdgeTable entry) */
PBF_MC_HD uint32_t pbf_mc_edge_mask(uint64_t row) {
  uint32_t m = 0;
  for (uint32_t t = 0; t < 16u; ++t) {
    const uint32_t e = (uint32_t)((row >> (4u * t)) & 0xFu);
    if (e == 0xFu) break;
    m |= 1u << e;
  }
  return m;
}

#endif /* PBF_MC_TABLES_H */
