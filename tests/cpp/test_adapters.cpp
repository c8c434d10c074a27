// C++ drop-in check: the reference's own synthetic-geometry tests (test/test-sfm.cpp sfm_triangulate_cube,
// the L-shape rig of sfm_refine_L_shape fed to sfm_solve) and the ImagePair / matcher entry points, written
// against the adapters in include/mvslam/ exactly the way the reference's tests call its functions.
#include <cstdio>
#include <algorithm>
#include <cstdlib>
#include <random>

#include <mvslam/image-pair.hpp>
#include <mvslam/pnp.hpp>
#include <mvslam/ba.hpp>

using namespace mvSLAM;

#define ASSERT_TRUE(c) do { if (!(c)) { std::printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, #c); return false; } } while (0)
#define ASSERT_EQUAL(a, b, tol) ASSERT_TRUE(std::fabs((a) - (b)) <= (tol))

static std::vector<Point3> rig(bool cube)
{
    std::vector<Point3> p;
    if (cube) { for (int x = -1; x <= 1; x += 2) for (int y = -1; y <= 1; y += 2) for (int z = -1; z <= 1; z += 2) p.emplace_back(x, y, z); }
    else p = {{1, 0, 0}, {0, 0, 0}, {0, 2, 0}, {1, 0, 3}, {0, 0, 3}, {0, 2, 3}, {0.5, 0, 1.5}, {0, 1, 1.5}};
    return p;
}

static void project(const std::vector<Point3> &P, double tx, std::vector<ImagePoint> &out)
{   // K = I, camera at (tx, 0, 0), no rotation
    out.clear();
    for (auto &p : P) out.emplace_back((p[0] - tx) / p[2], p[1] / p[2]);
}

static bool sfm_triangulate_cube()   // test/test-sfm.cpp:92-155
{
    auto P = rig(true);
    for (auto &p : P) { p[0] += 0.6; p[2] += 3.0; }
    std::vector<ImagePoint> x1, x2;
    project(P, 0.0, x1); project(P, 1.0, x2);
    std::vector<Point3> pts; std::vector<size_t> idx;
    sfm_triangulate(x1, x2, Matrix3Type::Identity(), SE3(), SE3(SO3(), Vector3Type(1, 0, 0)), pts, idx);
    ASSERT_TRUE(pts.size() == P.size());
    for (size_t i = 0; i < P.size(); ++i) for (int j = 0; j < 3; ++j) ASSERT_EQUAL(P[i][j], pts[i][j], 1e-3);
    return true;
}

static bool sfm_solve_L_shape()      // rig of test/test-sfm.cpp:173-177, solved with the own 8-point branch
{
    const double roll = 1.5, pitch = 0.7;
    const double cr = std::cos(roll), sr = std::sin(roll), cp = std::cos(pitch), sp = std::sin(pitch);
    auto P = rig(false);
    for (auto &p : P) {   // R = Ry(pitch) * Rx(roll), scale 0.5, t = (0.6, 0, 3)
        double x = 0.5 * p[0], y = 0.5 * p[1], z = 0.5 * p[2];
        double y1 = cr * y - sr * z, z1 = sr * y + cr * z;
        p = Point3(cp * x + sp * z1 + 0.6, y1, -sp * x + cp * z1 + 3.0);
    }
    std::vector<ImagePoint> x1, x2;
    project(P, 0.0, x1); project(P, 1.0, x2);
    Transformation pose; std::vector<Point3> pts; std::vector<size_t> idx;
    ASSERT_TRUE(sfm_solve(x1, x2, Matrix3Type::Identity(), pose, pts, idx));
    ASSERT_EQUAL(pose.translation()[0], 1.0, 1e-3); ASSERT_EQUAL(pose.translation()[1], 0.0, 1e-3); ASSERT_EQUAL(pose.translation()[2], 0.0, 1e-3);
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) ASSERT_EQUAL(pose.rotation().get_matrix()(i, j), i == j ? 1.0 : 0.0, 1e-3);
    ASSERT_TRUE(pts.size() == 8);
    for (size_t i = 0; i < 8; ++i) for (int j = 0; j < 3; ++j) ASSERT_EQUAL(P[i][j], pts[i][j], 1e-3);
    // too few points: false, outputs untouched
    x1.resize(5); x2.resize(5);
    Transformation keep = pose;
    ASSERT_TRUE(!sfm_solve(x1, x2, Matrix3Type::Identity(), pose, pts, idx));
    ASSERT_EQUAL(pose.translation()[0], keep.translation()[0], 0.0);
    return true;
}

static bool image_pair_and_matcher()
{
    std::mt19937 rng(7);
    const int n = 600;
    // scene + two cameras (K = [500 0 320; 0 500 240]), descriptors = random 256 bit, pair frame permuted with bit noise
    Matrix3Type K = Matrix3Type::Identity(); K(0, 0) = K(1, 1) = 500; K(0, 2) = 320; K(1, 2) = 240;
    std::uniform_real_distribution<double> ux(-3, 3), uz(4, 10);
    VisualFeatureConfig::DetectorResultType k1(n), k2(n);
    VisualFeatureConfig::ExtractorResultType d1(n * 32), d2(n * 32);
    std::vector<int> perm(n);
    for (int i = 0; i < n; ++i) perm[i] = i;
    std::shuffle(perm.begin(), perm.end(), rng);
    for (int i = 0; i < n; ++i) {
        double X = ux(rng), Y = ux(rng), Z = uz(rng);
        k1[i].pt.x = (float)(500 * X / Z + 320); k1[i].pt.y = (float)(500 * Y / Z + 240);
        int j = perm[i];
        k2[j].pt.x = (float)(500 * (X - 0.5) / Z + 320); k2[j].pt.y = (float)(500 * Y / Z + 240);
        for (int b = 0; b < 32; ++b) { d1[i * 32 + b] = (uint8_t)rng(); d2[j * 32 + b] = d1[i * 32 + b] ^ (uint8_t)(1u << (rng() % 8)) * (b % 11 == 0); }
    }
    auto f1 = std::make_shared<Frame>(); f1->id = 0; f1->visual_feature = VisualFeature(k1, d1, 640, 480);
    auto f2 = std::make_shared<Frame>(); f2->id = 1; f2->visual_feature = VisualFeature(k2, d2, 640, 480);
    auto matches = VisualFeature::match_visual_features(f1->visual_feature, f2->visual_feature, 10);
    ASSERT_TRUE(matches.size() == (size_t)n);
    for (size_t i = 0; i < matches.size(); ++i) {
        ASSERT_TRUE(perm[matches[i].trainIdx] == matches[i].queryIdx);
        if (i) ASSERT_TRUE(matches[i - 1].distance <= matches[i].distance);
    }
    auto both = VisualFeature::match_and_filter_visual_features(f1->visual_feature, f2->visual_feature, 10);
    ASSERT_TRUE(both.first.size() == (size_t)n && both.second.size() == (size_t)n);
    b200::ransac_defaults().n_hypotheses = 64;
    b200::ransac_defaults().max_error_sq = 1e-6;      // float keypoints: ~1e-5 px noise
    ImagePair ip(f1, f2, K, ImagePair::get_default_params());
    ASSERT_TRUE(ip.valid);
    ASSERT_TRUE(ip.match_inlier_count > (uint32_t)(0.9 * n));
    ASSERT_EQUAL(ip.T_pair_to_base.translation()[0], 1.0, 1e-3);     // camera 2 sits at +x; |t| = 1
    for (const auto &mp : ip.matched_points) ASSERT_TRUE(perm[mp.vf_idx_in_base] == (int)mp.vf_idx_in_pair);
    {   // ImagePair::refine (image-pair.cpp:176-237): two-view bundle adjustment keeps the solution, adds covariances
        ImagePair refined = ip;
        ASSERT_TRUE(refined.refine() && refined.refined);
        ASSERT_TRUE(refined.error >= 0 && refined.error < 1.0);
        ASSERT_EQUAL(refined.T_pair_to_base.translation()[0], 1.0, 2e-2);
        ASSERT_TRUE(refined.matched_points_covar.size() == refined.matched_points.size());
        for (int k = 0; k < 6; ++k) ASSERT_TRUE(refined.T_pair_to_base_covar(k, k) > 0);
    }
    auto batch = ImagePair::solve_batch({f1, f2}, {{0, 1}, {1, 0}}, K, ImagePair::get_default_params());
    ASSERT_TRUE(batch.size() == 2 && batch[0].valid && batch[1].valid);
    ASSERT_EQUAL(batch[1].T_pair_to_base.translation()[0], -1.0, 1e-3);
    return true;
}

// test/test-pnp.cpp:14-63 (pnp_solve_cube): cube rig at (0.6, 0, 3), camera at x = +1, K = I, tolerance 1e-3, no outliers
static bool pnp_solve_cube()
{
    std::vector<Point3> world;
    std::vector<ImagePoint> image;
    for (int x = -1; x <= 1; x += 2) for (int y = -1; y <= 1; y += 2) for (int z = -1; z <= 1; z += 2) {
        Point3 p(x + 0.6, y + 0.0, z + 3.0);
        world.push_back(p);
        image.emplace_back((p[0] - 1.0) / p[2], p[1] / p[2]);
    }
    Transformation pose;
    std::vector<size_t> inliers;
    ASSERT_TRUE(pnp_solve(world, image, Matrix3Type::Identity(), pose, inliers));
    ASSERT_TRUE(inliers.size() == world.size());
    ASSERT_EQUAL(pose.translation()[0], 1.0, 1e-3);
    ASSERT_EQUAL(pose.translation()[1], 0.0, 1e-3);
    ASSERT_EQUAL(pose.translation()[2], 0.0, 1e-3);
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) ASSERT_EQUAL(pose.rotation().get_matrix()(r, c), r == c ? 1.0 : 0.0, 1e-3);
    return true;
}

// test/test-sfm.cpp:157-290 (sfm_refine_L_shape): noisy observations, guesses and points; everything back within 0.025
static bool sfm_refine_L_shape()
{
    const double tol = 0.025, noise = 5e-3;
    std::mt19937 rng(3);
    std::normal_distribution<double> g(0.0, 1.0);
    const double raw[8][3] = {{1, 0, 0}, {0, 0, 0}, {0, 2, 0}, {1, 0, 3}, {0, 0, 3}, {0, 2, 3}, {0.5, 0, 1.5}, {0, 1, 1.5}};
    const double cy = std::cos(1.5), sy = std::sin(1.5), cp = std::cos(0.7), sp = std::sin(0.7);
    std::vector<Point3> X, Xg;
    std::vector<Point2Estimate> p1, p2;
    Matrix2Type C; C(0, 0) = C(1, 1) = noise * noise;
    for (auto &r : raw) {   // 0.5 * Rz(1.5) Ry(0.7) p + (0.6, 0, 3)
        const double a = cp * r[0] + sp * r[2], b = r[1], c = -sp * r[0] + cp * r[2];
        Point3 p(0.5 * (cy * a - sy * b) + 0.6, 0.5 * (sy * a + cy * b), 0.5 * c + 3.0);
        X.push_back(p);
        Xg.emplace_back(p[0] + 5e-3 * g(rng), p[1] + 5e-3 * g(rng), p[2] + 5e-3 * g(rng));
        Point2 a1, a2;
        a1[0] = p[0] / p[2] + noise * g(rng); a1[1] = p[1] / p[2] + noise * g(rng);
        a2[0] = (p[0] - 1.0) / p[2] + noise * g(rng); a2[1] = p[1] / p[2] + noise * g(rng);
        p1.emplace_back(a1, C); p2.emplace_back(a2, C);
    }
    Transformation guess(SO3(Matrix3Type::Identity()), Vector3Type(1.0 + 4e-3, -3e-3, 5e-3));
    TransformationEstimate pose;
    std::vector<Point3Estimate> pts;
    ScalarType err = -1;
    ASSERT_TRUE(sfm_refine(p1, p2, Matrix3Type::Identity(), guess, Xg, pose, pts, err));
    ASSERT_TRUE(err >= 0 && pts.size() == X.size());
    ASSERT_EQUAL(pose.mean().translation()[0], 1.0, tol);
    ASSERT_EQUAL(pose.mean().translation()[1], 0.0, tol);
    ASSERT_EQUAL(pose.mean().translation()[2], 0.0, tol);
    for (size_t i = 0; i < X.size(); ++i) for (int k = 0; k < 3; ++k) ASSERT_EQUAL(pts[i].mean()[k], X[i][k], tol);
    for (int k = 0; k < 6; ++k) ASSERT_TRUE(pose.covar()(k, k) > 0);
    // pnp_refine on the same scene: camera 2 from the noisy points and its observations (test-pnp.cpp:65-160)
    std::vector<Point3Estimate> world;
    Matrix3Type Cw; Cw(0, 0) = Cw(1, 1) = Cw(2, 2) = 25e-6;
    for (auto &p : Xg) world.emplace_back(p, Cw);
    TransformationEstimate cam;
    ASSERT_TRUE(pnp_refine(world, p2, Matrix3Type::Identity(), guess, cam, err));
    ASSERT_EQUAL(cam.mean().translation()[0], 1.0, tol);
    return true;
}

// VisualFeature::extract (visual-feature.hpp:14) on two renderings of the same rectangles shifted by 6 pixels:
// the extracted features match each other under that shift, like utility/test-visual-feature.cpp:25-35 shows by eye.
static bool extract_and_match()
{
    const int W = 320, H = 240, shift = 6;
    std::mt19937 rng(11);
    struct Box { int x, y, w, h, v; };
    std::vector<Box> boxes(60);
    for (auto &b : boxes) b = {(int)(rng() % W), (int)(rng() % H), 8 + (int)(rng() % 50), 8 + (int)(rng() % 40), (int)(rng() % 256)};
    auto render = [&](int dx) {
        std::vector<uint8_t> im((size_t)W * H, 90);
        for (const auto &b : boxes)
            for (int y = std::max(b.y, 0); y < std::min(b.y + b.h, H); ++y)
                for (int x = std::max(b.x + dx, 0); x < std::min(b.x + dx + b.w, W); ++x) im[(size_t)y * W + x] = (uint8_t)b.v;
        return im;
    };
    const std::vector<uint8_t> a = render(0), b = render(shift);
    VisualFeature f1 = VisualFeature::extract(ImageGrayscale(H, W, a.data()));
    VisualFeature f2 = VisualFeature::extract(ImageGrayscale(H, W, b.data()));
    ASSERT_TRUE(f1.valid() && f2.valid());
    ASSERT_TRUE(f1.size() > 100 && f1.size() <= 500 + 64);
    auto batch = VisualFeature::extract_batch({ImageGrayscale(H, W, a.data()), ImageGrayscale(H, W, b.data())});
    ASSERT_TRUE(batch.size() == 2 && batch[0].get_descriptors() == f1.get_descriptors() && batch[1].get_descriptors() == f2.get_descriptors());
    for (size_t i = 1; i < f1.size(); ++i) ASSERT_TRUE(f1.get_keypoints()[i - 1].octave <= f1.get_keypoints()[i].octave);
    auto matches = VisualFeature::match_visual_features(f1, f2, 30);
    ASSERT_TRUE(matches.size() > 50);
    size_t consistent = 0;
    for (const auto &m : matches) {
        const auto &p1 = f1.get_keypoints()[m.trainIdx].pt; const auto &p2 = f2.get_keypoints()[m.queryIdx].pt;
        consistent += std::fabs(p2.x - p1.x - shift) < 2.5f && std::fabs(p2.y - p1.y) < 2.5f;
    }
    ASSERT_TRUE(consistent * 10 >= matches.size() * 9);
    return true;
}

int main()
{
    int fails = 0;
    struct { const char *name; bool (*fn)(); } tests[] = {{"sfm_triangulate_cube", sfm_triangulate_cube},
                                                           {"sfm_solve_L_shape", sfm_solve_L_shape},
                                                           {"image_pair_and_matcher", image_pair_and_matcher},
                                                           {"extract_and_match", extract_and_match},
                                                           {"pnp_solve_cube", pnp_solve_cube},
                                                           {"sfm_refine_L_shape", sfm_refine_L_shape}};
    for (auto &t : tests) {
        bool ok = false;
        try { ok = t.fn(); } catch (const std::exception &e) { std::printf("exception: %s\n", e.what()); }
        std::printf("%s %s\n", ok ? "PASSED" : "FAILED", t.name);
        fails += !ok;
    }
    return fails;
}
