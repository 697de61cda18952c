"""INTEGRATION.md documents the body a maintainer of the reference puts into ImageAlignment::align.  This image has no
Eigen / Sophus / OpenCV headers, so that body had never met a compiler.  Here it is cut out of INTEGRATION.md verbatim and
compiled -- together with the real-type adapters of host/svo_types.hpp -- against minimal stand-ins that carry the real
libraries' signatures (tests/cpp/mock/), inside a harness that declares the members of the reference classes the body
touches (include/frame.hpp, include/feature.hpp, include/point.hpp, include/image_alignment.hpp).  The adapters also run:
SE3 <-> Sophus::SE3d and Mat8 <-> cv::Mat round trips."""
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "semi-direct-visual-odometry_b200")

HARNESS = r'''
#include <cmath>
#include <cstdio>
#include <memory>
#include <stdexcept>
#include <vector>
#include <Eigen/Core>
#include <sophus/se3.hpp>
#include <opencv2/core.hpp>
#include "svo_types.hpp"           // the adapters under test (sees the stand-in headers through __has_include)
#include "../../include/svo_b200.h"
#if !defined(SVO_HAVE_EIGEN) || !defined(SVO_HAVE_SOPHUS) || !defined(SVO_HAVE_OPENCV)
#error "the adapters were not compiled in"
#endif
// ---- the members of the reference classes the binding touches ----
struct Point { Eigen::Vector3d m_position; };                                              // include/point.hpp
struct Frame;
struct Feature {                                                                           // include/feature.hpp
    Eigen::Vector2d m_pixelPosition; Eigen::Vector3d m_bearingVec; std::shared_ptr< Point > m_point;
};
struct ImagePyramid { int m_slot = 0; int slot() const { return m_slot; } };               // (the slot replaces two vector<cv::Mat>)
struct Frame {                                                                             // include/frame.hpp
    std::vector< std::shared_ptr< Feature > > m_features; Sophus::SE3d m_absPose; ImagePyramid m_imagePyramid;
    std::shared_ptr< Frame > m_lastKeyframe;
    std::size_t numberObservation() const { return m_features.size(); }
};
struct Optimizer { uint32_t m_maxIteration = 20; };
struct ImageAlignment {                                                                    // include/image_alignment.hpp
    uint32_t m_patchSize = 5; int32_t m_minLevel = 0, m_maxLevel = 3; Optimizer m_optimizer;
    svo_ctx* m_gpu = nullptr; std::vector< svo_align_feature > m_packed;
    double align( std::shared_ptr< Frame >& refFrame, std::shared_ptr< Frame >& curFrame );
};
// ---- link-time stand-ins of the two C-ABI calls (this is a compile-and-run check of the HOST code, no GPU) ----
static svo_align_job g_job; static int g_n;
extern "C" svo_status svo_sparse_align( svo_ctx*, const svo_align_job* jobs, int, const svo_align_feature*, int n_feats,
                                        const svo_align_params*, svo_align_result* results, svo_align_level_stats* )
{
    g_job = jobs[ 0 ]; g_n = n_feats;
    for ( int i = 0; i < 7; i++ ) results->T_cur[ i ] = jobs[ 0 ].T_cur[ i ];
    results->T_cur[ 4 ] += 0.5; results->rmse = 3.25;
    return SVO_OK;
}
extern "C" const char* svo_last_error( const svo_ctx* ) { return ""; }
// ---- INTEGRATION.md, verbatim ----
@BODY@
int main()
{
    auto kf = std::make_shared< Frame >(); auto ref = std::make_shared< Frame >(); auto cur = std::make_shared< Frame >();
    ref->m_lastKeyframe = kf; ref->m_imagePyramid.m_slot = 3; kf->m_imagePyramid.m_slot = 1; cur->m_imagePyramid.m_slot = 4;
    for ( int i = 0; i < 3; i++ ) {
        auto f = std::make_shared< Feature >(); f->m_pixelPosition = Eigen::Vector2d( 10 + i, 20 ); f->m_bearingVec = Eigen::Vector3d( 0, 0, 1 );
        if ( i ) { f->m_point = std::make_shared< Point >(); f->m_point->m_position = Eigen::Vector3d( 1, 2, 3 + i ); }
        ( i < 2 ? ref : kf )->m_features.push_back( f );
    }
    cur->m_absPose = Sophus::SE3d( Eigen::Quaterniond( 1, 0, 0, 0 ), Eigen::Vector3d( 0.1, 0.2, 0.3 ) );
    ImageAlignment a;
    const double e = a.align( ref, cur );
    const auto p = cur->m_absPose.params();
    bool ok = e == 3.25 && g_n == 3 && g_job.ref_slot == 3 && g_job.kf_slot == 1 && g_job.cur_slot == 4 && g_job.n_ref == 2 && g_job.n_kf == 1
              && std::fabs( p[ 4 ] - 0.6 ) < 1e-15 && p[ 3 ] == 1.0 && a.m_packed[ 1 ].has_point && !a.m_packed[ 0 ].has_point && a.m_packed[ 2 ].point[ 2 ] == 5.0;
    // adapters
    const svo::SE3 T = svo::fromSophus( cur->m_absPose ); const Sophus::SE3d S = svo::toSophus( T );
    for ( int i = 0; i < 7; i++ ) ok = ok && S.params()[ i ] == cur->m_absPose.params()[ i ];
    ok = ok && svo::fromEigen( Eigen::Vector3d( 1, 2, 3 ) ).z() == 3 && svo::toEigen( svo::Vec2( 4, 5 ) ).y() == 5;
    unsigned char px[ 12 ] = { 1, 2, 3, 0, 4, 5, 6, 0, 7, 8, 9, 0 };
    const cv::Mat m( 3, 3, CV_8UC1, px, 4 );                                               // a strided cv::Mat
    const svo::Mat8 g = svo::fromCv( m ); const cv::Mat back = svo::toCv( g );
    ok = ok && g.at( 2, 1 ) == 8 && back.isContinuous() && back.ptr< uint8_t >( 1 )[ 2 ] == 6;
    std::printf( ok ? "INTEGRATION BINDING OK\n" : "MISMATCH\n" );
    return ok ? 0 : 1;
}
'''


def test_documented_binding_compiles_and_runs(tmp_path):
    md = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    blocks = re.findall(r"```cpp\n(.*?)```", md, flags=re.S)
    body = [b for b in blocks if "double ImageAlignment::align(" in b]
    assert len(body) == 1, "INTEGRATION.md must hold exactly one body of ImageAlignment::align"
    src = tmp_path / "binding.cpp"
    src.write_text(HARNESS.replace("@BODY@", body[0]))
    exe = tmp_path / "binding"
    subprocess.check_call(["g++", "-std=c++17", "-Wall", "-Wextra", "-Werror", "-Wno-unused-parameter", "-I", os.path.join(ROOT, "tests", "cpp", "mock"),
                           "-I", os.path.join(PKG, "host"), "-I", os.path.join(ROOT, "tests", "cpp"), str(src), "-o", str(exe)])
    out = subprocess.run([str(exe)], stdout=subprocess.PIPE, text=True)
    assert out.returncode == 0 and "INTEGRATION BINDING OK" in out.stdout, out.stdout
