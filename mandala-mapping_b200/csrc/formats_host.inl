/* formats_host.inl — the reference's persistence formats without PCL / Boost, part of m3dreg.cu.
 *
 * PCD.  The reference writes scans with pcl::io::savePCDFileBinary(path, pcl::PointCloud<PointXYZIRNLRGB>)
 * (src/gpu6DSLAM.cpp:41,88) and reads them with pcl::io::loadPCDFile (src/gpu6DSLAM.cpp:693).  PCL is a third-party
 * dependency that is not part of /root/reference (find_package(PCL), unpinned; 1.7.2 on the Ubuntu 16.04 / ROS kinetic the
 * tree targets).  Its published writer for a typed cloud (pcl/io/impl/pcd_io.hpp, PCDWriter::generateHeader<PointT> +
 * PCDWriter::writeBinary<PointT>) emits one header line per keyword for the REGISTERED fields only (padding is skipped) and
 * then the fields of every point packed back to back — 38 bytes per point for the fields registered in
 * include/custom_point_types.h:22-32, not the 40-byte struct.  That layout is what is written here; the reader maps fields
 * by NAME and honours SIZE / COUNT (so padded layouts with `_` fields, as the PCLPointCloud2 writer produces, load too).
 *
 * XML.  class data_model (include/data_model.hpp, src/data_model.cpp) is a boost::property_tree saved with
 * write_xml(..., xml_writer_make_settings<std::string>('\t', 1)): nested elements in insertion order, one per line, tab
 * indentation, text escaped.  Matrices are written with operator<<(float) — six significant digits — column by column. */
#include <string>
#include <vector>
#include <map>
#include <memory>
#include <cerrno>
#include <sys/stat.h>

namespace {

/* ---- PCD ------------------------------------------------------------------------------------------------------------- */
struct PcdField { std::string name; int size = 4; char type = 'F'; int count = 1; int offset = 0; };

struct PointFieldDesc { const char *name; int size; char type; size_t offset; };
const PointFieldDesc kPointFields[] = {      /* include/custom_point_types.h:22-32, in registration order */
	{"x", 4, 'F', offsetof(m3dreg_point, x)}, {"y", 4, 'F', offsetof(m3dreg_point, y)}, {"z", 4, 'F', offsetof(m3dreg_point, z)},
	{"intensity", 4, 'F', offsetof(m3dreg_point, intensity)}, {"ring", 2, 'U', offsetof(m3dreg_point, ring)},
	{"normal_x", 4, 'F', offsetof(m3dreg_point, normal_x)}, {"normal_y", 4, 'F', offsetof(m3dreg_point, normal_y)},
	{"normal_z", 4, 'F', offsetof(m3dreg_point, normal_z)}, {"label", 4, 'I', offsetof(m3dreg_point, label)},
	{"rgb", 4, 'F', offsetof(m3dreg_point, rgb)}};
constexpr int kPointFieldCount = (int)(sizeof(kPointFields) / sizeof(kPointFields[0]));

std::vector<std::string> split_ws(const std::string &line)
{
	std::vector<std::string> out;
	size_t i = 0;
	while (i < line.size()) {
		while (i < line.size() && isspace((unsigned char)line[i])) i++;
		size_t j = i;
		while (j < line.size() && !isspace((unsigned char)line[j])) j++;
		if (j > i) out.push_back(line.substr(i, j - i));
		i = j;
	}
	return out;
}

/* value of a field element as double, from its binary image */
double pcd_load_value(const unsigned char *p, int size, char type)
{
	if (type == 'F') { if (size == 4) { float v; memcpy(&v, p, 4); return v; } if (size == 8) { double v; memcpy(&v, p, 8); return v; } }
	if (type == 'U') { if (size == 1) return *p; if (size == 2) { uint16_t v; memcpy(&v, p, 2); return v; } if (size == 4) { uint32_t v; memcpy(&v, p, 4); return v; } }
	if (type == 'I') { if (size == 1) return (signed char)*p; if (size == 2) { int16_t v; memcpy(&v, p, 2); return v; } if (size == 4) { int32_t v; memcpy(&v, p, 4); return v; } }
	return 0.0;
}

void pcd_store_field(m3dreg_point &pt, int k, const unsigned char *bin, int size, char type, const char *ascii)
{
	const PointFieldDesc &d = kPointFields[k];
	unsigned char *dst = reinterpret_cast<unsigned char *>(&pt) + d.offset;
	if (bin && size == d.size && type == d.type) { memcpy(dst, bin, (size_t)d.size); return; }      /* same representation: bit copy */
	if (bin && d.name[0] == 'r' && d.name[1] == 'g' && size == 4) { memcpy(dst, bin, 4); return; }     /* rgb packed as U32 or F32: same bits */
	const double v = bin ? pcd_load_value(bin, size, type) : strtod(ascii, nullptr);
	if (d.type == 'F') { float f = (float)v; memcpy(dst, &f, 4); }
	else if (d.size == 2) { uint16_t u = (uint16_t)v; memcpy(dst, &u, 2); }
	else { int32_t i = (int32_t)v; memcpy(dst, &i, 4); }
}

} /* namespace */

extern "C" {

int m3dreg_pcd_write_binary(const char *path, const m3dreg_point *cloud, int n)
{
	if (!path || n < 0 || (n > 0 && !cloud)) return M3DREG_E_INVALID_ARG;
	FILE *f = fopen(path, "wb");
	if (!f) return M3DREG_E_IO;
	std::string names, sizes, types, counts;
	int rec = 0;
	for (int k = 0; k < kPointFieldCount; k++) {
		names += std::string(" ") + kPointFields[k].name;
		sizes += " " + std::to_string(kPointFields[k].size);
		types += std::string(" ") + kPointFields[k].type;
		counts += " 1";
		rec += kPointFields[k].size;
	}
	fprintf(f, "# .PCD v0.7 - Point Cloud Data file format\nVERSION 0.7\nFIELDS%s\nSIZE%s\nTYPE%s\nCOUNT%s\nWIDTH %d\nHEIGHT 1\n"
			"VIEWPOINT 0 0 0 1 0 0 0\nPOINTS %d\nDATA binary\n", names.c_str(), sizes.c_str(), types.c_str(), counts.c_str(), n, n);
	std::vector<unsigned char> buf((size_t)rec * 4096);
	bool ok = true;
	for (int base = 0; base < n && ok; base += 4096) {
		const int cnt = n - base < 4096 ? n - base : 4096;
		unsigned char *o = buf.data();
		for (int i = 0; i < cnt; i++) {
			const unsigned char *p = reinterpret_cast<const unsigned char *>(cloud + base + i);
			for (int k = 0; k < kPointFieldCount; k++) { memcpy(o, p + kPointFields[k].offset, (size_t)kPointFields[k].size); o += kPointFields[k].size; }
		}
		ok = fwrite(buf.data(), (size_t)rec, (size_t)cnt, f) == (size_t)cnt;
	}
	ok = (fclose(f) == 0) && ok;
	return ok ? 0 : M3DREG_E_IO;
}

int m3dreg_pcd_read(const char *path, m3dreg_point *out, int cap, int *n_out)
{
	if (!path || !n_out) return M3DREG_E_INVALID_ARG;
	*n_out = 0;
	FILE *f = fopen(path, "rb");
	if (!f) return M3DREG_E_IO;
	std::vector<PcdField> fields;
	long long points = -1, width = -1, height = 1;
	int data_kind = -1;      /* 0 ascii, 1 binary */
	char line[4096];
	while (fgets(line, sizeof(line), f)) {
		std::vector<std::string> t = split_ws(line);
		if (t.empty() || t[0][0] == '#') continue;
		const std::string &kw = t[0];
		if (kw == "FIELDS" || kw == "COLUMNS") { fields.resize(t.size() - 1); for (size_t i = 1; i < t.size(); i++) fields[i - 1].name = t[i]; }
		else if (kw == "SIZE") { for (size_t i = 1; i < t.size() && i - 1 < fields.size(); i++) fields[i - 1].size = atoi(t[i].c_str()); }
		else if (kw == "TYPE") { for (size_t i = 1; i < t.size() && i - 1 < fields.size(); i++) fields[i - 1].type = t[i][0]; }
		else if (kw == "COUNT") { for (size_t i = 1; i < t.size() && i - 1 < fields.size(); i++) fields[i - 1].count = atoi(t[i].c_str()); }
		else if (kw == "WIDTH" && t.size() > 1) width = atoll(t[1].c_str());
		else if (kw == "HEIGHT" && t.size() > 1) height = atoll(t[1].c_str());
		else if (kw == "POINTS" && t.size() > 1) points = atoll(t[1].c_str());
		else if (kw == "DATA" && t.size() > 1) { data_kind = t[1] == "binary" ? 1 : (t[1] == "ascii" ? 0 : 2); break; }
	}
	if (points < 0 && width >= 0) points = width * height;
	if (fields.empty() || points < 0 || points > 2147483647LL || data_kind < 0 || data_kind > 1) { fclose(f); return M3DREG_E_IO; }
	int rec = 0;
	std::vector<int> target(fields.size(), -1);
	bool have_xyz[3] = {false, false, false};
	for (size_t i = 0; i < fields.size(); i++) {
		PcdField &fd = fields[i];
		if (fd.count <= 0) fd.count = 1;
		if (fd.size <= 0 || fd.size > 8) { fclose(f); return M3DREG_E_IO; }
		fd.offset = rec;
		rec += fd.size * fd.count;
		for (int k = 0; k < kPointFieldCount; k++) if (fd.name == kPointFields[k].name && fd.count == 1) { target[i] = k; if (k < 3) have_xyz[k] = true; }
	}
	if (!(have_xyz[0] && have_xyz[1] && have_xyz[2])) { fclose(f); return M3DREG_E_IO; }
	*n_out = (int)points;
	if (!out) { fclose(f); return 0; }
	if (points > cap) { fclose(f); return M3DREG_E_SIZE_MISMATCH; }
	int rc = 0;
	if (data_kind == 1) {
		std::vector<unsigned char> buf((size_t)rec * 4096);
		for (long long base = 0; base < points && rc == 0; base += 4096) {
			const size_t cnt = (size_t)(points - base < 4096 ? points - base : 4096);
			if (fread(buf.data(), (size_t)rec, cnt, f) != cnt) { rc = M3DREG_E_IO; break; }
			for (size_t i = 0; i < cnt; i++) {
				m3dreg_point pt;
				memset(&pt, 0, sizeof(pt));
				const unsigned char *p = buf.data() + i * (size_t)rec;
				for (size_t k = 0; k < fields.size(); k++) if (target[k] >= 0) pcd_store_field(pt, target[k], p + fields[k].offset, fields[k].size, fields[k].type, nullptr);
				out[base + (long long)i] = pt;
			}
		}
	} else {
		for (long long i = 0; i < points && rc == 0; i++) {
			if (!fgets(line, sizeof(line), f)) { rc = M3DREG_E_IO; break; }
			std::vector<std::string> t = split_ws(line);
			m3dreg_point pt;
			memset(&pt, 0, sizeof(pt));
			size_t col = 0;
			for (size_t k = 0; k < fields.size(); k++) {
				if (col + (size_t)fields[k].count > t.size()) { rc = M3DREG_E_IO; break; }
				if (target[k] >= 0) pcd_store_field(pt, target[k], nullptr, 0, fields[k].type, t[col].c_str());
				col += (size_t)fields[k].count;
			}
			out[i] = pt;
		}
	}
	fclose(f);
	return rc;
}

} /* extern "C" */

/* ---- XML model ------------------------------------------------------------------------------------------------------- */
struct m3dreg_model {
	struct Node {
		std::string key, text;
		std::vector<std::unique_ptr<Node>> kids;      /* insertion order, as boost::property_tree keeps it */
		Node *find(const std::string &k) const { for (auto &c : kids) if (c->key == k) return c.get(); return nullptr; }
		Node *get_or_add(const std::string &k)
		{
			if (Node *n = find(k)) return n;
			kids.emplace_back(new Node());
			kids.back()->key = k;
			return kids.back().get();
		}
	};
	Node root;
	std::string xml_path;

	/* dotted path, as ptree::put / get_child_optional address nodes */
	Node *at(const std::string &path, bool create)
	{
		Node *n = &root;
		size_t i = 0;
		while (i <= path.size()) {
			size_t j = path.find('.', i);
			if (j == std::string::npos) j = path.size();
			const std::string k = path.substr(i, j - i);
			Node *c = create ? n->get_or_add(k) : n->find(k);
			if (!c) return nullptr;
			n = c;
			i = j + 1;
		}
		return n;
	}
	const Node *at(const std::string &path) const { return const_cast<m3dreg_model *>(this)->at(path, false); }
	void put(const std::string &path, const std::string &value) { at(path, true)->text = value; }
};

namespace {

std::string xml_escape(const std::string &s)
{
	std::string o;
	for (char ch : s) {
		switch (ch) {
		case '<': o += "&lt;"; break;
		case '>': o += "&gt;"; break;
		case '&': o += "&amp;"; break;
		case '"': o += "&quot;"; break;
		case '\'': o += "&apos;"; break;
		default: o += ch;
		}
	}
	return o;
}

std::string xml_unescape(const std::string &s)
{
	std::string o;
	for (size_t i = 0; i < s.size(); i++) {
		if (s[i] == '&') {
			const struct { const char *e; char c; } ents[] = {{"&lt;", '<'}, {"&gt;", '>'}, {"&amp;", '&'}, {"&quot;", '"'}, {"&apos;", '\''}};
			bool hit = false;
			for (auto &e : ents) { size_t l = strlen(e.e); if (s.compare(i, l, e.e) == 0) { o += e.c; i += l - 1; hit = true; break; } }
			if (hit) continue;
		}
		o += s[i];
	}
	return o;
}

/* boost::property_tree::xml_parser::write_xml_element with indent char '\t', count 1 */
void xml_write(FILE *f, const m3dreg_model::Node &n, int depth)
{
	const std::string ind((size_t)depth, '\t');
	if (n.kids.empty()) {
		if (n.text.empty()) fprintf(f, "%s<%s/>\n", ind.c_str(), n.key.c_str());
		else fprintf(f, "%s<%s>%s</%s>\n", ind.c_str(), n.key.c_str(), xml_escape(n.text).c_str(), n.key.c_str());
		return;
	}
	fprintf(f, "%s<%s>\n", ind.c_str(), n.key.c_str());
	if (!n.text.empty()) fprintf(f, "%s\t%s\n", ind.c_str(), xml_escape(n.text).c_str());
	for (auto &c : n.kids) xml_write(f, *c, depth + 1);
	fprintf(f, "%s</%s>\n", ind.c_str(), n.key.c_str());
}

/* the subset of XML property_tree's writer produces (and hand-edited files of the same shape): declaration, comments,
 * elements without attributes of interest, character data */
bool xml_parse(const std::string &s, m3dreg_model::Node &root)
{
	std::vector<m3dreg_model::Node *> stack{&root};
	size_t i = 0;
	while (i < s.size()) {
		if (s[i] != '<') {
			size_t j = s.find('<', i);
			if (j == std::string::npos) j = s.size();
			std::string txt = s.substr(i, j - i);
			size_t a = txt.find_first_not_of(" \t\r\n"), b = txt.find_last_not_of(" \t\r\n");
			if (a != std::string::npos && stack.size() > 1) stack.back()->text += xml_unescape(txt.substr(a, b - a + 1));
			i = j;
			continue;
		}
		if (s.compare(i, 4, "<!--") == 0) { size_t j = s.find("-->", i); if (j == std::string::npos) return false; i = j + 3; continue; }
		if (s.compare(i, 2, "<?") == 0) { size_t j = s.find("?>", i); if (j == std::string::npos) return false; i = j + 2; continue; }
		if (s.compare(i, 2, "<!") == 0) { size_t j = s.find('>', i); if (j == std::string::npos) return false; i = j + 1; continue; }
		size_t j = s.find('>', i);
		if (j == std::string::npos) return false;
		std::string tag = s.substr(i + 1, j - i - 1);
		i = j + 1;
		if (!tag.empty() && tag[0] == '/') {
			if (stack.size() <= 1 || stack.back()->key != tag.substr(1, tag.find_first_of(" \t\r\n", 1) - 1)) return false;
			stack.pop_back();
			continue;
		}
		const bool self = !tag.empty() && tag.back() == '/';
		if (self) tag.pop_back();
		const std::string name = tag.substr(0, tag.find_first_of(" \t\r\n"));
		if (name.empty()) return false;
		stack.back()->kids.emplace_back(new m3dreg_model::Node());
		stack.back()->kids.back()->key = name;
		if (!self) stack.push_back(stack.back()->kids.back().get());
	}
	return stack.size() == 1;
}

/* operator<<(std::ostream &, float) with the default format: %g, six significant digits */
std::string fmt_float(float v)
{
	char b[64];
	snprintf(b, sizeof(b), "%g", (double)v);
	return b;
}

int copy_out(const std::string &s, char *out, int cap)
{
	if (!out || cap <= 0) return (int)s.size();
	const size_t n = s.size() < (size_t)cap - 1 ? s.size() : (size_t)cap - 1;
	memcpy(out, s.data(), n);
	out[n] = 0;
	return (int)s.size();
}

std::string dir_of(const std::string &path)
{
	const size_t k = path.find_last_of('/');
	return k == std::string::npos ? std::string(".") : (k == 0 ? std::string("/") : path.substr(0, k));
}

} /* namespace */

extern "C" {

m3dreg_model *m3dreg_model_create(void) { return new (std::nothrow) m3dreg_model(); }
void m3dreg_model_destroy(m3dreg_model *m) { delete m; }

int m3dreg_model_load(m3dreg_model *m, const char *xml_path)
{
	if (!m || !xml_path) return M3DREG_E_INVALID_ARG;
	m->root.kids.clear();                                  /* pt_.clear() */
	m->xml_path = xml_path;
	FILE *f = fopen(xml_path, "rb");
	if (!f) return M3DREG_E_IO;
	std::string s;
	char buf[65536];
	size_t got;
	while ((got = fread(buf, 1, sizeof(buf), f)) > 0) s.append(buf, got);
	fclose(f);
	if (!xml_parse(s, m->root)) { m->root.kids.clear(); return M3DREG_E_IO; }
	return 0;
}

int m3dreg_model_save(const m3dreg_model *m, const char *xml_path)
{
	if (!m || !xml_path) return M3DREG_E_INVALID_ARG;
	FILE *f = fopen(xml_path, "wb");
	if (!f) return M3DREG_E_IO;
	fprintf(f, "<?xml version=\"1.0\" encoding=\"utf-8\"?>\n");
	for (auto &c : m->root.kids) xml_write(f, *c, 0);
	return fclose(f) == 0 ? 0 : M3DREG_E_IO;
}

void m3dreg_model_set_algorithm_name(m3dreg_model *m, const char *name) { if (m && name) m->put("Model.Algorithms.name", name); }
void m3dreg_model_set_dataset_path(m3dreg_model *m, const char *path) { if (m && path) m->put("Model.DatasetPath", path); }

int m3dreg_model_get_dataset_path(const m3dreg_model *m, char *out, int cap)
{
	if (!m) return M3DREG_E_INVALID_ARG;
	const m3dreg_model::Node *n = m->at("Model.DatasetPath");
	return copy_out(n ? n->text : std::string(), out, cap);
}

void m3dreg_model_set_affine(m3dreg_model *m, const char *scan_id, const float *a)
{
	if (!m || !scan_id || !a) return;
	const std::string base = std::string("Model.Transformations.") + scan_id + ".Affine.";
	m->put(base + "Type", "matrix4f");
	std::string data;
	for (int c = 0; c < 4; c++) for (int r = 0; r < 4; r++) data += fmt_float(a[r * 4 + c]) + " ";      /* column by column, data_model.cpp:143-147 */
	m->put(base + "Data", data);
}

int m3dreg_model_get_affine(const m3dreg_model *m, const char *scan_id, float *a)
{
	if (!m || !scan_id || !a) return M3DREG_E_INVALID_ARG;
	const std::string base = std::string("Model.Transformations.") + scan_id + ".Affine.";
	const m3dreg_model::Node *ty = m->at(base + "Type"), *da = m->at(base + "Data");
	if (!ty || !da) return M3DREG_E_BAD_SLOT;
	std::vector<std::string> t = split_ws(da->text);
	for (int k = 0; k < 16; k++) a[k] = (k % 5 == 0) ? 1.0f : 0.0f;
	if (ty->text == "matrix4f") {
		if (t.size() < 16) return M3DREG_E_IO;
		for (int c = 0; c < 4; c++) for (int r = 0; r < 4; r++) a[r * 4 + c] = strtof(t[(size_t)(c * 4 + r)].c_str(), nullptr);
		return 0;
	}
	if (ty->text == "Vector3f_Quaternionf") {              /* origin x y z, quaternion x y z w (data_model.cpp:75-87) */
		if (t.size() < 7) return M3DREG_E_IO;
		float v[7];
		for (int k = 0; k < 7; k++) v[k] = strtof(t[(size_t)k].c_str(), nullptr);
		const float x = v[3], y = v[4], z = v[5], w = v[6];
		a[0] = 1 - 2 * (y * y + z * z); a[1] = 2 * (x * y - w * z); a[2] = 2 * (x * z + w * y); a[3] = v[0];
		a[4] = 2 * (x * y + w * z); a[5] = 1 - 2 * (x * x + z * z); a[6] = 2 * (y * z - w * x); a[7] = v[1];
		a[8] = 2 * (x * z - w * y); a[9] = 2 * (y * z + w * x); a[10] = 1 - 2 * (x * x + y * y); a[11] = v[2];
		return 0;
	}
	return M3DREG_E_IO;
}

void m3dreg_model_set_cloud_name(m3dreg_model *m, const char *scan_id, const char *fn)
{
	if (m && scan_id && fn) m->put(std::string("Model.Transformations.") + scan_id + ".cloudname", fn);
}

int m3dreg_model_get_cloud_name(const m3dreg_model *m, const char *scan_id, char *out, int cap)
{
	if (!m || !scan_id) return M3DREG_E_INVALID_ARG;
	const m3dreg_model::Node *n = m->at(std::string("Model.Transformations.") + scan_id + ".cloudname");
	if (!n) return M3DREG_E_BAD_SLOT;
	return copy_out(n->text, out, cap);
}

int m3dreg_model_scan_count(const m3dreg_model *m)
{
	if (!m) return M3DREG_E_INVALID_ARG;
	const m3dreg_model::Node *n = m->at("Model.Transformations");
	return n ? (int)n->kids.size() : 0;
}

int m3dreg_model_scan_id(const m3dreg_model *m, int index, char *out, int cap)
{
	if (!m) return M3DREG_E_INVALID_ARG;
	const m3dreg_model::Node *n = m->at("Model.Transformations");
	if (!n || index < 0 || (size_t)index >= n->kids.size()) return M3DREG_E_BAD_SLOT;
	return copy_out(n->kids[(size_t)index]->key, out, cap);
}

int m3dreg_model_full_cloud_path(const m3dreg_model *m, const char *scan_id, char *out, int cap)
{
	if (!m || !scan_id) return M3DREG_E_INVALID_ARG;
	const m3dreg_model::Node *ds = m->at("Model.DatasetPath");
	const m3dreg_model::Node *cn = m->at(std::string("Model.Transformations.") + scan_id + ".cloudname");
	if (!cn) return M3DREG_E_BAD_SLOT;
	std::string p = dir_of(m->xml_path);
	if (ds && !ds->text.empty()) p += "/" + ds->text;
	p += "/" + cn->text;
	return copy_out(p, out, cap);
}

} /* extern "C" */
