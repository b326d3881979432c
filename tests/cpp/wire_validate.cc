// CPU-only check of the L1 validators restated in gyeeta_b200/csrc/gysk_wire.h (common/gy_comm_proto.cc:840-996): the accept /
// reject rules the reference applies before a batch reaches handle_l2_misc, on hand-built TCP_CONN_NOTIFY / AGGR_TASK_STATE_NOTIFY /
// LISTENER_STATE_NOTIFY batches. Prints one line per case; exit code 0 iff every case behaves as the reference's validate() does.
#include <cstdio>
#include <cstring>
#include <vector>

#include "gysk_wire.h"

using namespace gysk::wire;

static int g_fail = 0;
static void expect(const char *what, bool got, bool want)
{
	std::printf("%-58s %s\n", what, got == want ? "ok" : "MISMATCH");
	if (got != want) g_fail++;
}

template <typename T, typename LenField>
static std::vector<uint8_t> batch(int n, const std::vector<size_t> & tails, LenField T::*len)
{
	std::vector<uint8_t> buf;
	for (int i = 0; i < n; ++i) {
		T t;
		std::memset(&t, 0, sizeof(t));
		const size_t tail = tails[i % tails.size()], sz = sizeof(T) + tail, pad = (8 - sz % 8) % 8;
		t.*len = (LenField)tail; t.padding_len_ = (uint8_t)pad;
		const uint8_t *p = reinterpret_cast<const uint8_t *>(&t);
		buf.insert(buf.end(), p, p + sizeof(T));
		buf.insert(buf.end(), tail, (uint8_t)'x');		// string without its NUL: the validator forces it
		buf.insert(buf.end(), pad, (uint8_t)0);
	}
	return buf;
}

int main()
{
	auto tlen = [](const TCP_CONN_NOTIFY & t) -> size_t { return t.cli_cmdline_len_; };
	auto alen = [](const AGGR_TASK_STATE_NOTIFY & t) -> size_t { return t.issue_string_len_; };
	auto llen = [](const LISTENER_STATE_NOTIFY & t) -> size_t { return t.issue_string_len_; };

	{	// well-formed variable-stride batch; the trailing strings get their NUL in place (:866-868)
		auto b = batch<TCP_CONN_NOTIFY>(5, {0, 13, 40, 255, 8}, &TCP_CONN_NOTIFY::cli_cmdline_len_);
		auto *p = reinterpret_cast<TCP_CONN_NOTIFY *>(b.data());
		expect("tcp: 5 records, ragged command lines", validate_batch<TCP_CONN_NOTIFY>(p, 5, b.data() + b.size(), TCP_CONN_NOTIFY::MAX_NUM_CONNS, tlen), true);
		const uint8_t *second = b.data() + p->get_elem_size();
		const auto *q = reinterpret_cast<const TCP_CONN_NOTIFY *>(second);
		expect("tcp: string of record 2 NUL-forced at its last byte", second[sizeof(TCP_CONN_NOTIFY) + q->cli_cmdline_len_ - 1] == 0, true);
		expect("tcp: fewer records present than announced", validate_batch<TCP_CONN_NOTIFY>(p, 6, b.data() + b.size(), TCP_CONN_NOTIFY::MAX_NUM_CONNS, tlen), false);
		expect("tcp: last record cut short", validate_batch<TCP_CONN_NOTIFY>(p, 5, b.data() + b.size() - 8, TCP_CONN_NOTIFY::MAX_NUM_CONNS, tlen), false);
		expect("tcp: nevents above MAX_NUM_CONNS", validate_batch<TCP_CONN_NOTIFY>(p, 2049, b.data() + b.size(), TCP_CONN_NOTIFY::MAX_NUM_CONNS, tlen), false);
		expect("tcp: zero records", validate_batch<TCP_CONN_NOTIFY>(p, 0, b.data(), TCP_CONN_NOTIFY::MAX_NUM_CONNS, tlen), true);
	}
	{	// an element size that is not a multiple of 8 is rejected (:858-860)
		auto b = batch<TCP_CONN_NOTIFY>(2, {5}, &TCP_CONN_NOTIFY::cli_cmdline_len_);
		auto *p = reinterpret_cast<TCP_CONN_NOTIFY *>(b.data());
		p->padding_len_ = 0;
		expect("tcp: element size not a multiple of 8", validate_batch<TCP_CONN_NOTIFY>(p, 2, b.data() + b.size(), TCP_CONN_NOTIFY::MAX_NUM_CONNS, tlen), false);
	}
	{
		auto b = batch<AGGR_TASK_STATE_NOTIFY>(1200, {0, 31, 7}, &AGGR_TASK_STATE_NOTIFY::issue_string_len_);
		auto *p = reinterpret_cast<AGGR_TASK_STATE_NOTIFY *>(b.data());
		expect("task: 1200 records (the maximum)", validate_batch<AGGR_TASK_STATE_NOTIFY>(p, 1200, b.data() + b.size(), AGGR_TASK_STATE_NOTIFY::MAX_NUM_TASKS, alen), true);
		expect("task: 1201 announced", validate_batch<AGGR_TASK_STATE_NOTIFY>(p, 1201, b.data() + b.size(), AGGR_TASK_STATE_NOTIFY::MAX_NUM_TASKS, alen), false);
	}
	{
		auto b = batch<LISTENER_STATE_NOTIFY>(512, {0, 100, 254}, &LISTENER_STATE_NOTIFY::issue_string_len_);
		auto *p = reinterpret_cast<LISTENER_STATE_NOTIFY *>(b.data());
		expect("listener: 512 records (the maximum)", validate_batch<LISTENER_STATE_NOTIFY>(p, 512, b.data() + b.size(), LISTENER_STATE_NOTIFY::MAX_NUM_LISTENERS, llen), true);
		expect("listener: buffer ends inside record 1", validate_batch<LISTENER_STATE_NOTIFY>(p, 512, b.data() + 40, LISTENER_STATE_NOTIFY::MAX_NUM_LISTENERS, llen), false);
	}
	std::printf("failures: %d\n", g_fail);
	return g_fail ? 1 : 0;
}
